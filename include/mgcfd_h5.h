/* mgcfd_h5.h -- C-ABI of the from-scratch HDF5 subset reader / writer (mg-cfd-app-op2_b200/host/h5lite.hpp) that
 * stands in for the libhdf5 calls behind the reference's level-file and solution-file I/O:
 *   op_decl_set_hdf5_infer_size / op_decl_map_hdf5 / op_decl_dat_hdf5   (euler3d.cpp:248-327)  -> mgcfd_h5_open, _info, _read_*
 *   op_fetch_data_hdf5_file                                             (euler3d.cpp:564,740-770) -> mgcfd_h5_create, _add, _finish
 * Host code only (no CUDA, no torch types): built as libmgcfd_h5.so next to libmgcfd_b200.so. */
#ifndef MGCFD_H5_H
#define MGCFD_H5_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgcfd_h5_reader mgcfd_h5_reader;
typedef struct mgcfd_h5_writer mgcfd_h5_writer;

enum { MGCFD_H5_I32 = 0, MGCFD_H5_I64 = 1, MGCFD_H5_F32 = 2, MGCFD_H5_F64 = 3 };

/* 1 if the file carries an HDF5 superblock signature (at offset 0, 512, 1024, ...) */
int mgcfd_h5_is_hdf5(const char *path);

/* Open for reading: parses the superblock and every group.  NULL on error; the text is left in err (errlen bytes). */
mgcfd_h5_reader *mgcfd_h5_open(const char *path, char *err, int errlen);
void mgcfd_h5_close(mgcfd_h5_reader *r);
int mgcfd_h5_superblock_version(const mgcfd_h5_reader *r);
int mgcfd_h5_count(const mgcfd_h5_reader *r);
const char *mgcfd_h5_name(const mgcfd_h5_reader *r, int i);      /* "group/dataset" for nested groups */
/* type_class: 0 fixed-point, 1 floating-point, 3 string; layout: 0 compact, 1 contiguous, 2 chunked; dims: up to 8 */
int mgcfd_h5_info(const mgcfd_h5_reader *r, const char *name, int *type_class, int *elem_bytes, int *is_signed,
                  int *layout, int *rank, unsigned long long *dims);
/* Read a whole dataset, converted to native little-endian double / int32 (row-major).  0 on success. */
int mgcfd_h5_read_f64(const mgcfd_h5_reader *r, const char *name, double *out, char *err, int errlen);
int mgcfd_h5_read_i32(const mgcfd_h5_reader *r, const char *name, int *out, char *err, int errlen);
/* Attributes of a dataset: 0 on success, -1 if absent or of another kind */
int mgcfd_h5_attr_int(const mgcfd_h5_reader *r, const char *dataset, const char *attr, long long *out);
int mgcfd_h5_attr_str(const mgcfd_h5_reader *r, const char *dataset, const char *attr, char *out, int cap);

/* Write: datasets are borrowed until mgcfd_h5_finish, which writes the file (superblock v0, contiguous little-endian
 * datasets with OP2's "size" / "dim" / "type" attributes) and frees the writer. */
mgcfd_h5_writer *mgcfd_h5_create(const char *path);
int mgcfd_h5_add(mgcfd_h5_writer *w, const char *name, int dtype, int rank, const unsigned long long *dims, const void *data);
int mgcfd_h5_finish(mgcfd_h5_writer *w, char *err, int errlen);

#ifdef __cplusplus
}
#endif
#endif
