// TEST INFRASTRUCTURE -- robustness driver for the HDF5 reader (parpe_b200/csrc/h5min.h): walks every group of each
// file given on the command line and reads every dataset through every accessor.  Damaged files must end in an
// H5Error (any std::exception), never in an out-of-bounds access: the test builds this with ASan/UBSan.
#include <cstdio>
#include <exception>
#include <string>

#include "../../parpe_b200/csrc/h5min.h"
#include "../../parpe_b200/csrc/h5wmin.h"

static int walk(const pb200::H5File& f, const std::string& g, int depth) {
    int n = 0;
    if (depth > 8) return 0;
    for (auto& name : f.list(g)) {
        const std::string p = (g == "/" ? "" : g) + "/" + name;
        bool is_group = false;
        try {
            f.list(p);
            is_group = true;
        } catch (std::exception const&) {
        }
        if (is_group) {
            n += walk(f, p, depth + 1);
            continue;
        }
        try { f.info(p); ++n; } catch (std::exception const&) {}
        try { f.read_double(p); } catch (std::exception const&) {}
        try { f.read_int(p); } catch (std::exception const&) {}
        try { f.read_strings(p); } catch (std::exception const&) {}
        try { f.read_vlen_int(p); } catch (std::exception const&) {}
        try { f.read_bytes(p); } catch (std::exception const&) {}
    }
    try { f.attribute_names(g); } catch (std::exception const&) {}
    for (const char* a : {"atol", "rtol", "maxsteps", "newton_maxsteps"})
        try { if (f.has_attribute(g, a)) f.attribute(g, a); } catch (std::exception const&) {}
    return n;
}

/// --write <path>: a file with every kind of object the writer (h5wmin.h) emits -- nested and empty groups, a group
/// with more children than a default symbol node holds, doubles / int32 / int64 / strings, empty datasets,
/// attributes -- read back through the reader; both under the sanitizers.
static int write_and_read(const char* path) {
    pb200::H5Writer w;
    w.doubles("/a/b/c/values", {1.5, 2.5, 3.5, 4.5, 5.5, 6.5}, {2, 3});
    w.ints("/a/ints", {1, -2, 3}, {1, 3});
    w.int64s("/a/b/big", {(int64_t)1 << 40, -7}, {2});
    w.strings("/a/labels", {"o0i0c1", "", "a longer label"});
    w.doubles("/a/empty", {}, {0, 4});
    w.group("/nothing/in/here");
    w.attr_double("/a", "atol", 1e-12);
    for (int i = 0; i < 70; ++i) w.doubles("/many/" + std::to_string(i), {(double)i}, {1});
    w.save(path);
    pb200::H5File f(path);
    int bad = 0;
    bad += f.read_double("/a/b/c/values") != std::vector<double>{1.5, 2.5, 3.5, 4.5, 5.5, 6.5};
    bad += f.read_int("/a/ints") != std::vector<int64_t>{1, -2, 3};
    bad += f.read_int("/a/b/big") != std::vector<int64_t>{(int64_t)1 << 40, -7};
    bad += f.read_strings("/a/labels") != std::vector<std::string>{"o0i0c1", "", "a longer label"};
    bad += f.info("/a/empty").dims != std::vector<uint64_t>{0, 4} || !f.read_double("/a/empty").empty();
    bad += !f.list("/nothing/in/here").empty();
    bad += f.attribute("/a", "atol") != std::vector<double>{1e-12};
    bad += f.list("/many").size() != 70 || f.read_double("/many/69") != std::vector<double>{69.0};
    std::printf("WRITE-%s\n", bad ? "MISMATCH" : "OK");
    return bad;
}

int main(int argc, char** argv) {
    if (argc == 3 && std::string(argv[1]) == "--write") return write_and_read(argv[2]);
    int opened = 0, datasets = 0;
    for (int i = 1; i < argc; ++i) {
        try {
            pb200::H5File f(argv[i]);
            ++opened;
            datasets += walk(f, "/", 0);
        } catch (std::exception const&) {
        }
    }
    std::printf("ROBUST-OK files=%d opened=%d datasets=%d\n", argc - 1, opened, datasets);
    return 0;
}
