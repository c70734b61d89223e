// h5lite_capi.cpp -- the C-ABI of include/mgcfd_h5.h over h5lite.hpp (no exceptions cross the boundary)
#include "h5lite.hpp"
#include "mgcfd_h5.h"

struct mgcfd_h5_reader {
    h5lite::File file;
    explicit mgcfd_h5_reader(const char *p) : file(p) {}
};
struct mgcfd_h5_writer {
    h5lite::Writer w;
    explicit mgcfd_h5_writer(const char *p) : w(p) {}
};

static void set_err(char *err, int errlen, const char *msg)
{
    if (err && errlen > 0) {
        strncpy(err, msg, (size_t)errlen - 1);
        err[errlen - 1] = '\0';
    }
}

extern "C" {

int mgcfd_h5_is_hdf5(const char *path) { return path && h5lite::File::is_hdf5(path) ? 1 : 0; }

mgcfd_h5_reader *mgcfd_h5_open(const char *path, char *err, int errlen)
{
    try {
        return new mgcfd_h5_reader(path);
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return nullptr;
    }
}
void mgcfd_h5_close(mgcfd_h5_reader *r) { delete r; }
int mgcfd_h5_superblock_version(const mgcfd_h5_reader *r) { return r->file.superblock_version(); }
int mgcfd_h5_count(const mgcfd_h5_reader *r) { return (int)r->file.names().size(); }
const char *mgcfd_h5_name(const mgcfd_h5_reader *r, int i)
{
    return i >= 0 && i < (int)r->file.names().size() ? r->file.names()[i].c_str() : nullptr;
}
int mgcfd_h5_info(const mgcfd_h5_reader *r, const char *name, int *type_class, int *elem_bytes, int *is_signed, int *layout,
                  int *rank, unsigned long long *dims)
{
    if (!r->file.has(name)) return -1;
    const h5lite::DatasetInfo &d = r->file.info(name);
    if (type_class) *type_class = d.type.cls;
    if (elem_bytes) *elem_bytes = (int)d.type.size;
    if (is_signed) *is_signed = d.type.is_signed ? 1 : 0;
    if (layout) *layout = d.layout;
    if (rank) *rank = (int)d.dims.size();
    if (dims)
        for (size_t i = 0; i < d.dims.size() && i < 8; i++) dims[i] = d.dims[i];
    return 0;
}
int mgcfd_h5_read_f64(const mgcfd_h5_reader *r, const char *name, double *out, char *err, int errlen)
{
    try {
        r->file.read_f64(name, out);
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}
int mgcfd_h5_read_i32(const mgcfd_h5_reader *r, const char *name, int *out, char *err, int errlen)
{
    try {
        r->file.read_i32(name, out);
        return 0;
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        return -1;
    }
}
int mgcfd_h5_attr_int(const mgcfd_h5_reader *r, const char *dataset, const char *attr, long long *out)
{
    try {
        const auto &a = r->file.info(dataset).attrs;
        auto it = a.find(attr);
        if (it == a.end() || (it->second.type.cls != 0 && it->second.type.cls != 1) || it->second.raw.empty()) return -1;
        *out = it->second.as_int();
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
int mgcfd_h5_attr_str(const mgcfd_h5_reader *r, const char *dataset, const char *attr, char *out, int cap)
{
    try {
        const auto &a = r->file.info(dataset).attrs;
        auto it = a.find(attr);
        if (it == a.end() || it->second.type.cls != 3) return -1;
        set_err(out, cap, it->second.as_string().c_str());
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}

mgcfd_h5_writer *mgcfd_h5_create(const char *path) { return path ? new mgcfd_h5_writer(path) : nullptr; }
int mgcfd_h5_add(mgcfd_h5_writer *w, const char *name, int dtype, int rank, const unsigned long long *dims, const void *data)
{
    try {
        if (dtype < 0 || dtype > 3 || rank < 0 || rank > 8) return -1;
        std::vector<uint64_t> d(dims, dims + rank);
        w->w.add(name, (h5lite::DType)dtype, d, data);
        return 0;
    } catch (const std::exception &) {
        return -1;
    }
}
int mgcfd_h5_finish(mgcfd_h5_writer *w, char *err, int errlen)
{
    int rc = 0;
    try {
        w->w.close();
    } catch (const std::exception &e) {
        set_err(err, errlen, e.what());
        rc = -1;
    }
    delete w;
    return rc;
}

}  // extern "C"
