// TEST INFRASTRUCTURE -- robustness driver for the HDF5 reader (parpe_b200/csrc/h5min.h): walks every group of each
// file given on the command line and reads every dataset through every accessor.  Damaged files must end in an
// H5Error (any std::exception), never in an out-of-bounds access: the test builds this with ASan/UBSan.
#include <cstdio>
#include <exception>
#include <string>

#include "../../parpe_b200/csrc/h5min.h"

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

int main(int argc, char** argv) {
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
