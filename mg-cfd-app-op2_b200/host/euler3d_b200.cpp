// euler3d_b200.cpp -- native host driver: the reference's euler3d.cpp main() restated over the C-ABI of
// libmgcfd_b200 (include/mgcfd_b200.h).  Same command line (config.h:107-127), same input deck (io.h:28-205), same
// initialisation and V-cycle schedule (euler3d.cpp:413-441, 458-641), same -v validation (euler3d.cpp:658-718) and
// output naming (euler3d.cpp:720-779).  Level and solution files are HDF5 files with the reference's dataset names
// (read and written by h5lite.hpp, a from-scratch implementation of the HDF5 subset OP2's op_decl_*_hdf5 /
// op_fetch_data_hdf5_file use -- the image has no libhdf5) or MGCFDBIN containers holding the same datasets (see
// meshgen.py); the format is detected per file.  Outputs are .h5 when the deck is HDF5 or --hdf5 is given.
//
//   euler3d_b200 -i input.dat [-d dir] [-o prefix] [-g cycles] [-v] [-b] [-I n] [-m partitioner] [-r method]
//                [--renumber] [--output-variables] [--output-fluxes] [--output-step-factors]
//                [--gpus N] [--same-device] [--variant owner|gather|colour|atomic] [--exact] [--loopwise]
//                [--hdf5] [--check-deck]
#include <getopt.h>
#include <sys/time.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <limits>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "h5lite.hpp"
#include "mgcfd_b200.h"

namespace {

struct Dataset {
    int dtype = 0;                       // 0 float64, 1 int32
    std::vector<uint64_t> dims;
    std::vector<unsigned char> bytes;
    const double *f64() const { return reinterpret_cast<const double *>(bytes.data()); }
    const int *i32() const { return reinterpret_cast<const int *>(bytes.data()); }
};
typedef std::map<std::string, Dataset> Container;

// an HDF5 level / solution file: every numeric dataset, floating-point ones as float64 and integer ones as int32
bool read_hdf5(const std::string &path, Container &out, std::string &err)
{
    try {
        h5lite::File f(path);
        for (const std::string &name : f.names()) {
            const h5lite::DatasetInfo &info = f.info(name);
            if (info.type.cls != 0 && info.type.cls != 1) continue;
            Dataset d;
            d.dtype = info.type.cls == 1 ? 0 : 1;
            d.dims = info.dims;
            d.bytes.resize(info.count() * (d.dtype == 0 ? 8 : 4));
            if (d.dtype == 0) f.read_f64(name, reinterpret_cast<double *>(d.bytes.data()));
            else f.read_i32(name, reinterpret_cast<int32_t *>(d.bytes.data()));
            out[name] = std::move(d);
        }
        return true;
    } catch (const std::exception &e) {
        err = e.what();
        return false;
    }
}

bool read_container(const std::string &path, Container &out, std::string &err)
{
    if (h5lite::File::is_hdf5(path)) return read_hdf5(path, out, err);
    std::ifstream f(path, std::ios::binary);
    if (!f) { err = "cannot open " + path; return false; }
    char magic[8];
    uint32_t version = 0, n = 0;
    f.read(magic, 8);
    f.read(reinterpret_cast<char *>(&version), 4);
    f.read(reinterpret_cast<char *>(&n), 4);
    if (!f || memcmp(magic, "MGCFDBIN", 8) != 0) {
        err = path + ": neither an HDF5 file nor an MGCFDBIN container";
        return false;
    }
    f.seekg(0, std::ios::end);
    const uint64_t file_size = (uint64_t)f.tellg();            // every length field below is bounded by it
    f.seekg(16, std::ios::beg);
    if (n > 4096) { err = path + ": implausible dataset count"; return false; }
    for (uint32_t i = 0; i < n; i++) {
        uint32_t len = 0, dtype = 0, ndim = 0;
        f.read(reinterpret_cast<char *>(&len), 4);
        if (!f || len > 4096 || len > file_size) { err = path + ": corrupt dataset name length"; return false; }
        std::string name(len, '\0');
        f.read(&name[0], len);
        f.read(reinterpret_cast<char *>(&dtype), 4);
        f.read(reinterpret_cast<char *>(&ndim), 4);
        if (!f || dtype > 1 || ndim < 1 || ndim > 8) { err = path + ": corrupt header of dataset " + name; return false; }
        Dataset d;
        d.dtype = (int)dtype;
        d.dims.resize(ndim);
        f.read(reinterpret_cast<char *>(d.dims.data()), 8 * ndim);
        uint64_t nbytes = 0;
        f.read(reinterpret_cast<char *>(&nbytes), 8);
        uint64_t count = 1;
        for (uint64_t dim : d.dims) { if (dim > file_size) { count = ~0ull; break; } count *= dim; if (count > file_size) break; }
        if (!f || nbytes > file_size || count * (dtype == 0 ? 8 : 4) != nbytes) { err = path + ": dataset " + name + " has inconsistent dims / size"; return false; }
        f.seekg((8 - f.tellg() % 8) % 8, std::ios::cur);
        d.bytes.resize(nbytes);
        f.read(reinterpret_cast<char *>(d.bytes.data()), (std::streamsize)nbytes);
        if (!f) { err = path + ": truncated dataset " + name; return false; }
        out[name] = std::move(d);
    }
    return true;
}

bool g_hdf5_output = false;       // outputs as .h5 (op_fetch_data_hdf5_file, euler3d.cpp:740-770) instead of .mgb

bool write_container(std::string path, const std::string &name, const double *data, uint64_t rows, uint64_t cols)
{
    if (g_hdf5_output) {
        if (path.size() > 4 && path.compare(path.size() - 4, 4, ".mgb") == 0) path.replace(path.size() - 4, 4, ".h5");
        try {
            h5lite::Writer w(path);
            w.add(name, h5lite::DType::F64, cols > 1 ? std::vector<uint64_t>{rows, cols} : std::vector<uint64_t>{rows}, data);
            w.close();
            return true;
        } catch (const std::exception &e) {
            fprintf(stderr, "%s\n", e.what());
            return false;
        }
    }
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    uint32_t version = 1, n = 1, len = (uint32_t)name.size(), dtype = 0, ndim = cols > 1 ? 2 : 1;
    uint64_t dims[2] = {rows, cols}, nbytes = rows * cols * 8;
    f.write("MGCFDBIN", 8);
    f.write(reinterpret_cast<char *>(&version), 4);
    f.write(reinterpret_cast<char *>(&n), 4);
    f.write(reinterpret_cast<char *>(&len), 4);
    f.write(name.data(), len);
    f.write(reinterpret_cast<char *>(&dtype), 4);
    f.write(reinterpret_cast<char *>(&ndim), 4);
    f.write(reinterpret_cast<char *>(dims), 8 * ndim);
    f.write(reinterpret_cast<char *>(&nbytes), 8);
    static const char zeros[8] = {0};
    f.write(zeros, (8 - f.tellp() % 8) % 8);
    f.write(reinterpret_cast<const char *>(data), (std::streamsize)nbytes);
    return (bool)f;
}

std::string trim(const std::string &s)
{
    size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}

struct Deck {
    int size = 0, levels = 0, base = 1, mesh_name = -1;     // base_array_index defaults to 1 (euler3d.cpp:92)
    std::vector<std::string> files;
};

// io.h:28-205: "key = value" lines, '#' comments, a [levels] section of "<index> = <file>" lines
bool read_input_dat(const std::string &path, Deck &d, std::string &err)
{
    std::ifstream f(path);
    if (!f) { err = "Error: Could not open input file '" + path + "'"; return false; }
    std::string line;
    bool have_size = false, have_levels = false, have_name = false, have_files = false;
    while (std::getline(f, line)) {
        if (!line.empty() && line[0] == '#') continue;
        if (!line.empty() && line[0] == '[') {
            std::string sec = trim(line);
            int count = sec == "[levels]" ? d.levels : (sec == "[mg_mapping]" ? d.levels - 1 : 0);
            if (sec == "[levels]") {
                if (!have_levels) { err = "Need to know number of levels before parsing level filenames"; return false; }
                d.files.assign(d.levels, "");
                have_files = true;
            }
            for (int i = 0; i < count; i++) {
                if (!std::getline(f, line)) { err = "Have reached EOF before reading all filenames"; return false; }
                size_t eq = line.find('=');
                if (eq == std::string::npos) { err = "Was expecting a key-value pair following " + sec; return false; }
                int idx = atoi(trim(line.substr(0, eq)).c_str());
                if (sec == "[levels]" && idx >= 0 && idx < d.levels) d.files[idx] = trim(line.substr(eq + 1));
            }
            continue;
        }
        size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        std::string key = trim(line.substr(0, eq)), value = trim(line.substr(eq + 1));
        if (key == "size") { d.size = atoi(value.c_str()); have_size = true; }
        else if (key == "num_levels") { d.levels = atoi(value.c_str()); have_levels = true; }
        else if (key == "base_array_index") d.base = atoi(value.c_str());
        else if (key == "mesh_name") {
            // const.h:48-51
            if (value == "fvcorr") d.mesh_name = 0;
            else if (value == "la_cascade") d.mesh_name = 1;
            else if (value == "rotor37") d.mesh_name = 2;
            else if (value == "m6wing") d.mesh_name = 3;
            else { err = "Unknown mesh_name '" + value + "'"; return false; }
            have_name = true;
        }
    }
    if (!have_size) { err = "size not present"; return false; }
    if (!have_levels) { err = "number of levels not present"; return false; }
    if (!have_name) { err = "mesh name not present"; return false; }
    if (!have_files) { err = "mesh filenames not present"; return false; }
    return true;
}

double wall()
{
    timeval t;
    gettimeofday(&t, nullptr);
    return t.tv_sec + 1e-6 * t.tv_usec;
}

// io.h:296-417: the CSV files the reference's run-scripts aggregate (aggregate-output-data.py).  Rows are appended;
// the header is written when the file is new or empty.  <prefix>P=<rank>.PerfData.csv and <prefix>P=<rank>.FileIoTimes.csv
std::string perf_path(const std::string &prefix, int rank, const char *what)
{
    std::string p = prefix;
    if (p.length() > 1 && p[p.size() - 1] != '/' && p[p.size() - 1] != '.') p += ".";
    return p + "P=" + std::to_string(rank) + "." + what + ".csv";
}
bool csv_needs_header(const std::string &path)
{
    std::ifstream f(path);
    return !f || f.peek() == std::ifstream::traits_type::eof();
}
void dump_perf_data(const std::string &prefix, int rank, const std::string &partitioner, const std::vector<double> &flux_seconds,
                    const std::vector<long long> &flux_iters)
{
    const std::string path = perf_path(prefix, rank, "PerfData");
    const bool header = csv_needs_header(path);
    std::ofstream out(path, std::ios_base::app);
    if (header) out << "rank,partitioner,kernel,level,computeTime,syncTime,iters" << std::endl;
    for (size_t l = 0; l < flux_iters.size(); l++)
        out << rank << ',' << partitioner << ",compute_flux_edge_kernel," << l << ',' << flux_seconds[l] << ',' << 0.0 << ','
            << flux_iters[l] << std::endl;
}
void dump_file_io_perf_data(const std::string &prefix, int rank, const std::string &partitioner, int write_interval, int n_writes,
                            double io_seconds, double walltime)
{
    const std::string path = perf_path(prefix, rank, "FileIoTimes");
    const bool header = csv_needs_header(path);
    std::ofstream out(path, std::ios_base::app);
    if (header) out << "rank,partitioner,level,writeInterval,numberOfWrites,fileIoTime,wallTime" << std::endl;
    out << rank << ',' << partitioner << ',' << 0 << ',' << write_interval << ',' << n_writes << ',' << io_seconds << ',' << walltime
        << std::endl;      // base level only, as the reference
}

struct Config {                               // config.h:64-103, defaults :129-165
    std::string input_file, input_dir, prefix, variant = "owner";
    int cycles = 25, flow_interval = 0, gpus = 1;
    bool validate = false, mem_bound = false, renumber = true, exact = false, loopwise = false, same_device = false;
    bool hdf5 = false, check_deck = false, timers = false;
    std::string partitioner, partitioner_method;   // -m / -r (config.h:203-240)
    int out_vars = 0, out_fluxes = 0, out_sf = 0;
};

// op_partition's method when given (geom | kway | geomkway), else the library name, else the default
std::string partitioner_name(const Config &conf)
{
    return !conf.partitioner_method.empty() ? conf.partitioner_method : (!conf.partitioner.empty() ? conf.partitioner : std::string("geom"));
}

#define CHECK(call)                                                                                  \
    do {                                                                                             \
        int rc_ = (call);                                                                            \
        if (rc_ != MGCFD_OK) {                                                                       \
            fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, mgcfd_last_error(ctx));             \
            return 1;                                                                                \
        }                                                                                            \
    } while (0)

}  // namespace

int main(int argc, char **argv)
{
    Config conf;
    static option long_opts[] = {
        {"help", no_argument, nullptr, 'h'}, {"config-filepath", required_argument, nullptr, 'c'},
        {"legacy-mode", no_argument, nullptr, 'l'}, {"input-file", required_argument, nullptr, 'i'},
        {"input-directory", required_argument, nullptr, 'd'}, {"papi-config-file", required_argument, nullptr, 'p'},
        {"output-file-prefix", required_argument, nullptr, 'o'}, {"num-cycles", required_argument, nullptr, 'g'},
        {"partitioner", required_argument, nullptr, 'm'}, {"partitioner-method", required_argument, nullptr, 'r'},
        {"renumber", no_argument, nullptr, 'n'}, {"validate", no_argument, nullptr, 'v'},
        {"measure-mem-bound", no_argument, nullptr, 'b'}, {"output-variables", no_argument, &conf.out_vars, 1},
        {"output-fluxes", no_argument, &conf.out_fluxes, 1}, {"output-step-factors", no_argument, &conf.out_sf, 1},
        {"output-flow-interval", required_argument, nullptr, 'I'}, {"gpus", required_argument, nullptr, 1001},
        {"variant", required_argument, nullptr, 1002}, {"exact", no_argument, nullptr, 1003},
        {"loopwise", no_argument, nullptr, 1004}, {"same-device", no_argument, nullptr, 1005},
        {"hdf5", no_argument, nullptr, 1006}, {"check-deck", no_argument, nullptr, 1007},
        {"timers", no_argument, nullptr, 1008}, {nullptr, 0, nullptr, 0}};
    int opt;
    while ((opt = getopt_long(argc, argv, "hc:li:d:p:o:g:m:r:vbI:", long_opts, nullptr)) != -1) {
        switch (opt) {
        case 'i': conf.input_file = optarg; break;
        case 'd': conf.input_dir = optarg; break;
        case 'o': conf.prefix = optarg; break;
        case 'g': conf.cycles = atoi(optarg); break;
        case 'v': conf.validate = true; break;
        case 'b': conf.mem_bound = true; break;
        case 'I': conf.flow_interval = atoi(optarg); break;
        case 'n': conf.renumber = true; break;
        case 'm': conf.partitioner = optarg; break;          // block | random | parmetis | ptscotch | inertial
        case 'r': conf.partitioner_method = optarg; break;   // geom | kway | geomkway
        case 'c': case 'p': break;                           // config file / PAPI selections: accepted, not used
        case 'l': fprintf(stderr, "legacy mode (renumbered dataset names) is not supported\n"); return 1;
        case 1001: conf.gpus = atoi(optarg); break;
        case 1002: conf.variant = optarg; break;
        case 1003: conf.exact = true; break;
        case 1004: conf.loopwise = true; break;
        case 1005: conf.same_device = true; break;       // all ranks of --gpus N on device 0 (tests on a one-GPU box)
        case 1006: conf.hdf5 = true; break;              // write outputs as HDF5 even for an MGCFDBIN deck
        case 1007: conf.check_deck = true; break;        // load the deck, print what was read, exit (no GPU needed)
        case 1008: conf.timers = true; break;            // per-loop device timers -> PerfData / op2_performance_data.csv
        case 0: break;
        case 'h':
        default:
            printf("usage: %s -i input.dat [-d dir] [-o prefix] [-g cycles] [-v] [-b] [-I n] [--output-variables] "
                   "[--output-fluxes] [--output-step-factors] [--gpus N] [--variant owner|gather|colour|atomic] "
                   "[--exact] [--loopwise] [--hdf5] [--check-deck] [--timers]\n", argv[0]);
            return opt == 'h' ? 0 : 1;
        }
    }
    if (conf.input_file.empty()) { printf("ERROR: input_file not set\n"); return 1; }
    const std::string dir = conf.input_dir.empty() ? "" : conf.input_dir + "/";
    Deck deck;
    std::string err;
    if (!read_input_dat(dir + conf.input_file, deck, err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
    const int levels = deck.levels;

    printf("-----------------------------------------------------\nLoading level files ...\n");
    std::vector<Container> files(levels);
    std::vector<mgcfd_level_host> lv(levels);
    for (int i = 0; i < levels; i++) {
        printf("Loading level %d / %d\n", i + 1, levels);
        if (!read_container(dir + deck.files[i], files[i], err)) { fprintf(stderr, "%s\n", err.c_str()); return 1; }
        Container &c = files[i];
        for (const char *need : {"node_coordinates", "edge-->node", "edge_weights", "bnd_node-->node", "bnd_node-->group",
                                 "bnd_node_weights"})
            if (!c.count(need)) { fprintf(stderr, "%s: dataset %s missing\n", deck.files[i].c_str(), need); return 1; }
        // op_decl_map_hdf5 / op_decl_dat_hdf5 reject datasets whose size, dim or type do not match the declaring set
        // (euler3d.cpp:262-312): the same checks before any pointer is taken
        auto shape_ok = [&](const char *name, int dtype, uint64_t rows, uint64_t cols) {
            const Dataset &d = c[name];
            const bool ok = d.dtype == dtype && !d.dims.empty() && d.dims[0] == rows &&
                            ((d.dims.size() == 1 && cols == 1) || (d.dims.size() == 2 && d.dims[1] == cols)) &&
                            d.bytes.size() == rows * cols * (dtype == 0 ? 8u : 4u);
            if (!ok) fprintf(stderr, "%s: dataset %s does not have the expected type / shape [%llu x %llu]\n", deck.files[i].c_str(), name,
                             (unsigned long long)rows, (unsigned long long)cols);
            return ok;
        };
        for (const char *need : {"node_coordinates", "edge-->node", "bnd_node-->node"})
            if (c[need].dims.empty() || c[need].dims[0] > 0x7fffffffull) { fprintf(stderr, "%s: dataset %s has no usable first dimension\n", deck.files[i].c_str(), need); return 1; }
        const uint64_t nn = c["node_coordinates"].dims[0], ne = c["edge-->node"].dims[0], nb = c["bnd_node-->node"].dims[0];
        if (!shape_ok("node_coordinates", 0, nn, 3) || !shape_ok("edge-->node", 1, ne, 2) || !shape_ok("edge_weights", 0, ne, 3) ||
            !shape_ok("bnd_node-->node", 1, nb, 1) || !shape_ok("bnd_node-->group", 1, nb, 1) || !shape_ok("bnd_node_weights", 0, nb, 3))
            return 1;
        if (i + 1 < levels && c.count("node-->mg_node") && !shape_ok("node-->mg_node", 1, nn, 1)) return 1;
        mgcfd_level_host &h = lv[i];
        memset(&h, 0, sizeof(h));
        h.n_nodes = h.n_owned_nodes = (int)c["node_coordinates"].dims[0];
        h.n_edges = (int)c["edge-->node"].dims[0];
        h.n_bnd_nodes = (int)c["bnd_node-->node"].dims[0];
        h.node_coordinates = c["node_coordinates"].f64();
        h.edge_to_node = c["edge-->node"].i32();
        h.edge_weights = c["edge_weights"].f64();
        h.bnd_node_to_node = c["bnd_node-->node"].i32();
        h.bnd_node_to_group = c["bnd_node-->group"].i32();
        h.bnd_node_weights = c["bnd_node_weights"].f64();
        if (i + 1 < levels) {
            if (!c.count("node-->mg_node")) { fprintf(stderr, "%s: dataset node-->mg_node missing\n", deck.files[i].c_str()); return 1; }
            h.node_to_mg_node = c["node-->mg_node"].i32();
        }
    }
    g_hdf5_output = conf.hdf5 || h5lite::File::is_hdf5(dir + deck.files[0]);
    if (conf.check_deck) {
        // what the loader got out of every level file: sizes and order-sensitive checksums
        for (int i = 0; i < levels; i++) {
            printf("level %d: format=%s nodes=%d edges=%d bnd_nodes=%d\n", i,
                   h5lite::File::is_hdf5(dir + deck.files[i]) ? "hdf5" : "mgcfdbin", lv[i].n_nodes, lv[i].n_edges, lv[i].n_bnd_nodes);
            for (auto &kv : files[i]) {
                const Dataset &d = kv.second;
                long double sum = 0;
                const size_t n = d.bytes.size() / (d.dtype == 0 ? 8 : 4);
                for (size_t k = 0; k < n; k++) sum += (long double)(k % 97 + 1) * (d.dtype == 0 ? (long double)d.f64()[k] : (long double)d.i32()[k]);
                printf("  %-20s %s rank=%zu checksum=%.12Le\n", kv.first.c_str(), d.dtype == 0 ? "f64" : "i32", d.dims.size(), sum);
            }
        }
        if (!conf.prefix.empty()) {
            // the perf CSVs of a run that did nothing (the writers run without a GPU)
            dump_perf_data(conf.prefix, 0, partitioner_name(conf), std::vector<double>(levels, 0.0), std::vector<long long>(levels, 0));
            dump_file_io_perf_data(conf.prefix, 0, partitioner_name(conf), conf.flow_interval, 0, 0.0, 0.0);
        }
        return 0;
    }
    // -v: solution.variables.L<l>.cycles=<g> with dataset p_variables_result_L<l> (euler3d.cpp:314-335)
    std::vector<Container> solution(levels);
    std::vector<const double *> variables_correct(levels, nullptr);
    if (conf.validate)
        for (int i = 0; i < levels; i++) {
            std::string stem = dir + "solution.variables.L" + std::to_string(i) + ".cycles=" + std::to_string(conf.cycles);
            std::string p = stem + ".h5";                    // the reference's name (euler3d.cpp:320); else the container
            if (access(p.c_str(), R_OK) != 0) p = stem + ".mgb";
            std::string name = "p_variables_result_L" + std::to_string(i);
            if (read_container(p, solution[i], err) && solution[i].count(name)) variables_correct[i] = solution[i][name].f64();
            else printf("Cannot find level %d solution file: %s\n", i, p.c_str());
        }

    // ---- contexts: one per GPU (domain decomposition when --gpus > 1), euler3d.cpp:340-409
    const int P = conf.gpus;
    mgcfd_options o;
    mgcfd_default_options(&o);
    o.renumber = conf.renumber;
    o.exact_arith = conf.exact;
    o.measure_mem_bound = conf.mem_bound ? 1 : 0;                  // -b: unstructured_stream_kernel after every RK stage (:518-525)
    o.flux_variant = conf.variant == "atomic" ? MGCFD_FLUX_ATOMIC : conf.variant == "colour" ? MGCFD_FLUX_COLOUR
                   : conf.variant == "gather" ? MGCFD_FLUX_GATHER : MGCFD_FLUX_OWNER;
    std::vector<mgcfd_ctx *> R(P, nullptr);
    std::vector<mgcfd_local_mesh *> LM(P, nullptr);
    std::vector<std::vector<int>> part(levels);
    mgcfd_ctx *ctx = nullptr;
    mgcfd_consts consts;
    mgcfd_compute_farfield_consts(&consts);                       // euler3d.cpp:157-189
    consts.mesh_name = deck.mesh_name;
    if (P > 1) {
        printf("-----------------------------------------------------\nPartitioning ...\n");
        std::vector<const int *> pp(levels);
        // op_partition(lib, method, ...) at euler3d.cpp:340-375: the method wins when given (geom | kway | geomkway),
        // else the library name decides (parmetis / ptscotch are k-way partitioners, inertial is geometric)
        std::string method = partitioner_name(conf);
        printf("partitioner: %s\n", method.c_str());
        for (int l = 0; l < levels; l++) {
            part[l].resize(lv[l].n_nodes);
            int rc = l == 0 ? mgcfd_partition_graph(lv[0].n_nodes, lv[0].node_coordinates, lv[0].n_edges, lv[0].edge_to_node,
                                                    deck.base, P, method.c_str(), part[0].data())
                            : mgcfd_partition_coarse(lv[l - 1].n_nodes, part[l - 1].data(), lv[l - 1].node_to_mg_node, deck.base,
                                                     lv[l].n_nodes, lv[l].n_edges, lv[l].edge_to_node, lv[l].node_coordinates,
                                                     part[l].data());
            if (rc) { fprintf(stderr, "partitioning failed\n"); return 1; }
            pp[l] = part[l].data();
        }
        for (int r = 0; r < P; r++)
            if (mgcfd_local_mesh_build(levels, lv.data(), deck.base, pp.data(), r, P, &LM[r])) { fprintf(stderr, "local mesh failed\n"); return 1; }
        printf("PARTITIONING COMPLETE\n");
    }
    for (int r = 0; r < P; r++) {
        o.rank = r;
        o.n_ranks = P;
        if (mgcfd_create(&R[r], conf.same_device ? 0 : r, levels, &o) != MGCFD_OK) { fprintf(stderr, "mgcfd_create: %s\n", mgcfd_last_error(nullptr)); return 1; }
        ctx = R[r];
        CHECK(mgcfd_decl_consts(ctx, &consts));                  // op_decl_const x7, :232-238
        for (int i = 0; i < levels; i++) {
            if (P > 1) CHECK(mgcfd_decl_level(ctx, i, mgcfd_local_mesh_level(LM[r], i), 0));
            else CHECK(mgcfd_decl_level(ctx, i, &lv[i], deck.base));
        }
        CHECK(mgcfd_plan(ctx));
        for (int i = 0; i < levels; i++) {                       // :413-432
            CHECK(mgcfd_loop_initialize_variables(ctx, i));
            CHECK(mgcfd_loop_zero_fluxes(ctx, i));
            CHECK(mgcfd_loop_zero_volumes(ctx, i));
            CHECK(mgcfd_loop_calculate_cell_volumes(ctx, i));
        }
        for (int l = 0; l < levels; l++) {                       // :436-441
            CHECK(mgcfd_loop_dampen_ewt_edges(ctx, l));
            CHECK(mgcfd_loop_dampen_ewt_bnd(ctx, l));
        }
    }
    ctx = R[0];

    if (conf.timers)      // device timers per call site: inside graph replay on one GPU (mode 3), launch by launch otherwise
        for (int r = 0; r < P; r++) CHECK(mgcfd_timers_enable(R[r], (P > 1 || conf.loopwise) ? 1 : 3));
    printf("-----------------------------------------------------\nCompute beginning\n");
    double t1 = wall(), file_io_seconds = 0.0;
    int n_file_io_writes = 0;
    if (P > 1 || !conf.loopwise) {
        // device-driven schedule (host checks of :480 and :544 deferred to the end of the run)
        // -I <n> (euler3d.cpp:552-571): the run is cut into pieces of n cycles and the level-0 flow is written after each
        const int piece = conf.flow_interval > 0 ? conf.flow_interval : conf.cycles;
        for (int done = 0; done < conf.cycles;) {
            const int k = std::min(piece, conf.cycles - done);
            for (int i = 0; i < k; i++) printf("Performing MG cycle %d / %d\n", done + i + 1, conf.cycles);
            int rc = P > 1 ? mgcfd_group_run_cycles(R.data(), P, k) : mgcfd_run_cycles(ctx, k);
            if (rc == MGCFD_ERR_MIN_DT) { printf("Fatal error during 'step factor' calculation\n"); return 1; }
            if (rc == MGCFD_ERR_BAD_VALS) { printf("Bad variable values detected, aborting\n"); return 1; }
            if (rc) { fprintf(stderr, "run failed (%d): %s\n", rc, mgcfd_last_error(ctx)); return 1; }
            done += k;
            if (conf.flow_interval > 0 && done % conf.flow_interval == 0) {
                std::vector<double> v((size_t)lv[0].n_nodes * 5, 0.0);
                if (P == 1) {
                    CHECK(mgcfd_fetch_dat(ctx, 0, "variables", v.data()));
                } else {
                    for (int r = 0; r < P; r++) {
                        const mgcfd_level_host *h = mgcfd_local_mesh_level(LM[r], 0);
                        std::vector<double> loc((size_t)h->n_nodes * 5);
                        CHECK(mgcfd_fetch_dat(R[r], 0, "variables", loc.data()));
                        for (int i = 0; i < h->n_owned_nodes; i++)
                            for (int d = 0; d < 5; d++) v[(size_t)h->global_node_id[i] * 5 + d] = loc[(size_t)i * 5 + d];
                    }
                }
                const double w0 = wall();
                write_container(conf.prefix + "variables.L0.cycle=" + std::to_string(done) + ".mgb", "p_variables", v.data(),
                                lv[0].n_nodes, 5);
                file_io_seconds += wall() - w0;
                n_file_io_writes++;
            }
        }
    } else {
        // euler3d.cpp:458-641, call site by call site
        int level = 0, mg_dir = 0, i = 0, bad_val_count = 0;
        double rms = 0.0, min_dt;
        while (i < conf.cycles) {
            if (level == 0) printf("Performing MG cycle %d / %d\n", i + 1, conf.cycles);
            CHECK(mgcfd_loop_copy_double(ctx, level));
            CHECK(mgcfd_loop_calculate_dt(ctx, level));
            min_dt = std::numeric_limits<double>::max();
            CHECK(mgcfd_loop_get_min_dt(ctx, level, &min_dt));
            if (min_dt < 0.0f) { printf("Fatal error during 'step factor' calculation, min_dt = %.5e\n", min_dt); return 1; }
            CHECK(mgcfd_loop_compute_step_factor(ctx, level, &min_dt));
            for (int rkCycle = 0; rkCycle < MGCFD_RK; rkCycle++) {
                CHECK(mgcfd_loop_compute_flux_edge(ctx, level));
                CHECK(mgcfd_loop_compute_bnd_node_flux(ctx, level));
                CHECK(mgcfd_loop_time_step(ctx, level, &rkCycle));
                if (conf.mem_bound) CHECK(mgcfd_loop_unstructured_stream(ctx, level));
            }
            CHECK(mgcfd_loop_residual(ctx, level));
            if (level == 0) {
                rms = 0.0;
                CHECK(mgcfd_loop_calc_rms(ctx, level, &rms));
                rms = sqrt(rms / double(lv[level].n_nodes));
                bad_val_count = 0;
                CHECK(mgcfd_loop_count_bad_vals(ctx, level, &bad_val_count));
                if (bad_val_count > 0) { printf("Bad variable values detected, aborting\n"); return 1; }
            }
            if (conf.flow_interval > 0 && ((i + 1) % conf.flow_interval) == 0 && level == 0) {      // :552-571
                std::vector<double> v((size_t)lv[0].n_nodes * 5);
                CHECK(mgcfd_fetch_dat(ctx, 0, "variables", v.data()));
                const double w0 = wall();
                write_container(conf.prefix + "variables.L0.cycle=" + std::to_string(i + 1) + ".mgb", "p_variables", v.data(),
                                lv[0].n_nodes, 5);
                file_io_seconds += wall() - w0;
                n_file_io_writes++;
            }
            if (levels <= 1) {
                i++;
            } else if (mg_dir == 0) {
                level++;
                CHECK(mgcfd_loop_up_pre(ctx, level));
                CHECK(mgcfd_loop_up(ctx, level));
                CHECK(mgcfd_loop_up_post(ctx, level));
                if (level == levels - 1) mg_dir = 1;
            } else {
                level--;
                CHECK(mgcfd_loop_down(ctx, level));
                if (level == 0) { mg_dir = 0; i++; }
            }
        }
    }
    for (int r = 0; r < P; r++) mgcfd_sync(R[r]);
    const double walltime = wall() - t1;
    printf("\nCompute complete\n");
    printf("Max total runtime = %f\n", walltime);

    // assemble file-order results (ranks return [owned | halo] in their local order)
    std::vector<std::vector<double>> vars(levels);
    auto fetch_all = [&](const char *dat, int l, int dim, std::vector<double> &out) -> int {
        out.assign((size_t)lv[l].n_nodes * dim, 0.0);
        if (P == 1) return mgcfd_fetch_dat(R[0], l, dat, out.data());
        for (int r = 0; r < P; r++) {
            const mgcfd_level_host *h = mgcfd_local_mesh_level(LM[r], l);
            std::vector<double> loc((size_t)h->n_nodes * dim);
            int rc = mgcfd_fetch_dat(R[r], l, dat, loc.data());
            if (rc) return rc;
            for (int i = 0; i < h->n_owned_nodes; i++)
                for (int d = 0; d < dim; d++) out[(size_t)h->global_node_id[i] * dim + d] = loc[(size_t)i * dim + d];
        }
        return MGCFD_OK;
    };

    if (conf.validate) {                                                                     // :658-718
        printf("-----------------------------------------------------\n");
        printf("Looking for NaN and infinity values ...");
        bool failed = false;
        for (int l = 0; l < levels && !failed; l++) {
            CHECK(fetch_all("variables", l, 5, vars[l]));
            int bad = 0;
            for (double v : vars[l]) bad += (std::isnan(v) || std::isinf(v)) ? 1 : 0;
            if (bad > 0) { printf("\nValue check of MG level %d failed: %d bad values detected\n", l, bad); failed = true; }
        }
        if (!failed) {
            printf(" None found\nValidating result against solution ...");
            bool validation_failed = false;
            for (int l = 0; l < levels; l++) {
                if (!variables_correct[l]) { printf("\n- Do not have solution for level %d, cannot validate\n", l); validation_failed = true; continue; }
                int count = 0;
                if (P == 1) {
                    CHECK(mgcfd_validate_level(ctx, l, variables_correct[l], &count));        // identify_differences + count_non_zeros on the device
                } else {
                    for (size_t k = 0; k < vars[l].size(); k++) {                            // validation.h:65-88
                        double tol = fabs(variables_correct[l][k] * 10.0e-8);
                        if (tol < 3.0e-19) tol = 3.0e-19;
                        if (fabs(vars[l][k] - variables_correct[l][k]) > tol) count++;
                    }
                }
                if (count > lv[l].n_nodes / 5000) {
                    validation_failed = true;
                    printf("\nValidation of MG level %d failed: %d incorrect values in 'variables' array\n", l, count);
                    break;
                }
            }
            printf(validation_failed ? "Validation failed\n" : " Result correct\nValidation passed\n");
        }
    }
    if (conf.out_vars || conf.out_fluxes || conf.out_sf) {                                    // :720-779
        printf("-----------------------------------------------------\nWriting out data...\n");
        const double w0 = wall();
        n_file_io_writes += levels * (conf.out_vars + conf.out_fluxes + conf.out_sf);
        for (int l = 0; l < levels; l++) {
            std::string suffix = ".L" + std::to_string(l) + ".cycles=" + std::to_string(conf.cycles) + ".mgb";
            std::vector<double> buf;
            if (conf.out_sf) { CHECK(fetch_all("step_factors", l, 1, buf)); write_container(conf.prefix + "step_factors" + suffix, "p_step_factors_result_L" + std::to_string(l), buf.data(), lv[l].n_nodes, 1); }
            if (conf.out_fluxes) { CHECK(fetch_all("fluxes", l, 5, buf)); write_container(conf.prefix + "fluxes" + suffix, "p_fluxes_result_L" + std::to_string(l), buf.data(), lv[l].n_nodes, 5); }
            if (conf.out_vars) { CHECK(fetch_all("variables", l, 5, buf)); write_container(conf.prefix + "variables" + suffix, "p_variables_result_L" + std::to_string(l), buf.data(), lv[l].n_nodes, 5); }
        }
        file_io_seconds += wall() - w0;
    }
    // io.h:296-417 / euler3d.cpp:780-819: per-rank flux-kernel rows and the base-level file I/O row (next to the other
    // outputs; with no -o prefix they go into the input directory instead of the working directory)
    {
        const std::string csv_prefix = conf.prefix.empty() ? dir : conf.prefix;
        std::vector<int> visits(levels, 0);                       // level visits of one V-cycle (euler3d.cpp:573-640)
        if (levels == 1) visits[0] = 1;
        else for (int l = 0; l < levels; l++) visits[l] = (l == 0 || l == levels - 1) ? 1 : 2;
        for (int r = 0; r < P; r++) {
            std::vector<long long> iters(levels);
            for (int l = 0; l < levels; l++) {
                const int n_edges = P == 1 ? lv[l].n_edges : mgcfd_local_mesh_level(LM[r], l)->n_edges;
                iters[l] = (long long)n_edges * MGCFD_RK * visits[l] * conf.cycles;
            }
            // compute_flux_edge_kernel time per level (--timers): the fused Runge-Kutta stage is the launch that holds it
            std::vector<double> secs(levels, 0.0);
            if (conf.timers)
                for (int l = 0; l < levels; l++)
                    for (const char *name : {"rk_stage", "compute_flux_edge"}) {
                        double ms = 0.0;
                        long long calls = 0, elems = 0;
                        if (mgcfd_timers_get(R[r], name, l, &ms, &calls, &elems) == MGCFD_OK) secs[l] += ms * 1e-3;
                    }
            dump_perf_data(csv_prefix, r, partitioner_name(conf), secs, iters);
        }
        // op2_performance_data.csv (what op_timings_to_csv writes, euler3d.cpp:651-656; read by run-scripts/aggregate-output-data.py
        // :41,:71 for `nranks` and the per-loop times): one row per rank and loop; times are zero without --timers
        {
            const std::string path = csv_prefix + "op2_performance_data.csv";
            std::ofstream out(path);
            out << "rank,thread,nranks,nthreads,count,total time,plan time,mpi time,GB used,GB total,kernel name" << std::endl;
            static const char *const loops[] = {"visit_begin", "copy_double", "calculate_dt", "get_min_dt", "compute_step_factor", "rk_stage",
                                                "compute_flux_edge", "compute_bnd_node_flux", "time_step", "unstructured_stream", "residual",
                                                "calc_rms", "count_bad_vals", "restrict", "up_pre", "up", "up_post", "down", "min_exchange",
                                                "halo_exchange"};
            for (int r = 0; r < P; r++) {
                bool any = false;
                for (const char *name : loops) {
                    double ms = 0.0;
                    long long calls = 0, elems = 0;
                    if (!conf.timers || mgcfd_timers_get(R[r], name, -1, &ms, &calls, &elems) != MGCFD_OK || calls == 0) continue;
                    out << r << ",0," << P << ",1," << calls << ',' << ms * 1e-3 << ",0,0,0,0," << name << std::endl;
                    any = true;
                }
                if (!any) out << r << ",0," << P << ",1," << conf.cycles << ',' << walltime << ",0,0,0,0,mgcfd_run_cycles" << std::endl;
            }
        }
        dump_file_io_perf_data(csv_prefix, 0, partitioner_name(conf), conf.flow_interval, n_file_io_writes, file_io_seconds, walltime);
    }
    printf("-----------------------------------------------------\nWinding down\n");
    for (int r = 0; r < P; r++) {
        mgcfd_destroy(R[r]);
        if (LM[r]) mgcfd_local_mesh_free(LM[r]);
    }
    return 0;
}
