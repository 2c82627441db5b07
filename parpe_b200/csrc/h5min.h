// h5min.h -- minimal read-only HDF5 parser for parPE's parameter-estimation input files.
//
// What it replaces in the reference: the HDF5 C/C++ library calls behind
// parpe::MultiConditionDataProviderHDF5 (src/parpeamici/multiConditionDataProvider.cpp:42-60 opens the
// file and records the dataset paths; hdf5Misc.cpp's hdf5Read2DDoubleHyperslab / hdf5Read1DIntegerHyperslab
// etc. read them; deps/AMICI/src/hdf5.cpp reads the /amiciOptions attributes).  No HDF5 library exists in
// this image, and the hot path needs only "give me dataset X as doubles / ints" and "attribute Y of
// group Z" once at pb200_create time -- so this is a from-the-format-specification reader of exactly the
// subset h5py / the HDF5 library write for such files (doc/hdf5.md:1-172, python/parpe/hdf5_pe_input.py):
//
//   superblock version 0-3; version-1 object headers with continuation blocks; old-style groups (symbol
//   table message -> version-1 B-tree -> symbol nodes -> local heap); dataspace messages v1/v2; fixed-point
//   and IEEE floating-point datatypes (little endian, 1/2/4/8 bytes); data layout message v3: compact,
//   contiguous, and chunked (version-1 chunk B-tree) with the deflate and shuffle filters (zlib resolved at
//   run time); attribute messages v1-v3 with scalar / simple dataspaces.
//
// Everything else (version-2 object headers and "dense" new-style groups of libver='latest' files,
// variable-length strings, compound types, external storage, big-endian data) is reported as an error,
// never guessed.  Pinned by reading the reference tree's own HDF5 file
// (deps/AMICI/tests/cpp/testOptions.h5, written by the HDF5 library) in tests/test_h5_reader.py.
#ifndef PB200_H5MIN_H
#define PB200_H5MIN_H

#include <dlfcn.h>
#include <stdint.h>

#include <cstdio>
#include <cstring>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

namespace pb200 {

class H5Error : public std::runtime_error {
  public:
    explicit H5Error(const std::string& m) : std::runtime_error("hdf5: " + m) {}
};

class H5File {
  public:
    struct Dataset {
        int rank = 0;
        std::vector<uint64_t> dims;
        int type_class = -1;  // 0 fixed point, 1 floating point, 3 string (fixed length), 9 variable-length sequence
        int type_size = 0;
        bool type_signed = true;
        // variable-length data (class 9): sequences of integers (the element type below) or strings
        int base_class = -1, base_size = 0;
        bool base_signed = true;
        bool vlen_string = false;
        uint64_t count() const {  // (saturates instead of wrapping: a damaged dataspace must not look small)
            uint64_t n = 1;
            for (auto d : dims)
                if (__builtin_mul_overflow(n, d, &n)) return ~(uint64_t)0;
            return n;
        }
    };

    explicit H5File(const std::string& path) {
        FILE* f = std::fopen(path.c_str(), "rb");
        if (!f) throw H5Error("cannot open " + path);
        std::fseek(f, 0, SEEK_END);
        const long n = std::ftell(f);
        std::fseek(f, 0, SEEK_SET);
        buf_.resize(n > 0 ? (size_t)n : 0);
        if (n > 0 && std::fread(buf_.data(), 1, (size_t)n, f) != (size_t)n) {
            std::fclose(f);
            throw H5Error("short read of " + path);
        }
        std::fclose(f);
        parse_superblock();
    }

    bool exists(const std::string& path) const {
        uint64_t a;
        return locate(path, a);
    }

    /// names of the children of a group, in the file's (sorted) order
    std::vector<std::string> list(const std::string& group) const {
        uint64_t a;
        if (!locate(group, a)) throw H5Error("no such group: " + group);
        std::vector<std::string> out;
        for (auto& kv : children(a)) out.push_back(kv.first);
        return out;
    }

    Dataset info(const std::string& path) const {
        uint64_t a;
        if (!locate(path, a)) throw H5Error("no such dataset: " + path);
        DsRaw r = dataset_raw(a, path);
        return r.d;
    }

    /// whole dataset converted to double (row-major, as stored)
    std::vector<double> read_double(const std::string& path) const {
        uint64_t a;
        if (!locate(path, a)) throw H5Error("no such dataset: " + path);
        DsRaw r = dataset_raw(a, path);
        std::vector<uint8_t> raw = read_raw(r, path);
        std::vector<double> out(r.d.count());
        convert(raw.data(), r.d, out.data(), out.size(), path);
        return out;
    }
    /// the elements' bytes as stored (fixed-length strings, any fixed-size type)
    std::vector<uint8_t> read_bytes(const std::string& path) const {
        uint64_t a;
        if (!locate(path, a)) throw H5Error("no such dataset: " + path);
        DsRaw r = dataset_raw(a, path);
        if (r.d.type_class == 9) throw H5Error(path + ": variable-length data has no raw form");
        return read_raw(r, path);
    }
    std::vector<int64_t> read_int(const std::string& path) const {
        uint64_t a;
        if (!locate(path, a)) throw H5Error("no such dataset: " + path);
        DsRaw r = dataset_raw(a, path);
        if (r.d.type_class != 0) throw H5Error(path + ": not an integer dataset");
        std::vector<uint8_t> raw = read_raw(r, path);
        std::vector<int64_t> out(r.d.count());
        for (size_t i = 0; i < out.size(); ++i) out[i] = get_int(raw.data() + i * r.d.type_size, r.d.type_size, r.d.type_signed);
        return out;
    }

    /// ragged integer dataset: a 1-D dataset of variable-length integer sequences (h5py.vlen_dtype(int32), what
    /// parPE writes for /fixedParameters/reinitializationIndices) -> one vector per element.  Each element is stored
    /// as (length, global-heap collection address, object index); the sequences live in global heap collections.
    std::vector<std::vector<int64_t>> read_vlen_int(const std::string& path) const {
        uint64_t a;
        if (!locate(path, a)) throw H5Error("no such dataset: " + path);
        DsRaw r = dataset_raw(a, path);
        if (r.d.type_class != 9 || r.d.vlen_string || r.d.base_class != 0)
            throw H5Error(path + ": not a dataset of variable-length integer sequences");
        if (r.d.type_size != (int)(4 + so_ + 4)) throw H5Error(path + ": unexpected size of a variable-length element");
        std::vector<uint8_t> raw = read_raw(r, path);
        std::vector<std::vector<int64_t>> out(r.d.count());
        for (size_t i = 0; i < out.size(); ++i) {
            const uint8_t* p = raw.data() + i * r.d.type_size;
            const uint64_t len = rd(p, 4);
            const uint64_t coll = rd(p + 4, so_);
            const uint64_t index = rd(p + 4 + so_, 4);
            if (len == 0) continue;
            if (coll == UNDEF || coll == 0) throw H5Error(path + ": variable-length element without heap object");
            size_t nbytes = 0;
            const uint8_t* data = global_heap_object(coll, index, nbytes, path);
            if (nbytes < len * (uint64_t)r.d.base_size) throw H5Error(path + ": heap object shorter than the sequence");
            out[i].resize(len);
            for (uint64_t e = 0; e < len; ++e) out[i][e] = get_int(data + e * r.d.base_size, r.d.base_size, r.d.base_signed);
        }
        return out;
    }

    /// dataset of strings: fixed-length (class 3, null / space padded) or variable-length (h5py's default for str:
    /// class 9 with the string bit; the characters live in global heap collections)
    std::vector<std::string> read_strings(const std::string& path) const {
        uint64_t a;
        if (!locate(path, a)) throw H5Error("no such dataset: " + path);
        DsRaw r = dataset_raw(a, path);
        std::vector<uint8_t> raw = read_raw(r, path);
        std::vector<std::string> out(r.d.count());
        if (r.d.type_class == 3) {
            for (size_t i = 0; i < out.size(); ++i) {
                const char* p = reinterpret_cast<const char*>(raw.data() + i * r.d.type_size);
                size_t len = strnlen(p, (size_t)r.d.type_size);
                while (len > 0 && p[len - 1] == ' ') --len;  // (space-padded strings)
                out[i].assign(p, len);
            }
        } else if (r.d.type_class == 9 && r.d.vlen_string) {
            if (r.d.type_size != (int)(4 + so_ + 4)) throw H5Error(path + ": unexpected size of a variable-length element");
            for (size_t i = 0; i < out.size(); ++i) {
                const uint8_t* p = raw.data() + i * r.d.type_size;
                const uint64_t len = rd(p, 4), coll = rd(p + 4, so_), index = rd(p + 4 + so_, 4);
                if (len == 0 || coll == UNDEF || coll == 0) continue;
                size_t nbytes = 0;
                const uint8_t* data = global_heap_object(coll, index, nbytes, path);
                if (nbytes < len) throw H5Error(path + ": heap object shorter than the string");
                out[i].assign(reinterpret_cast<const char*>(data), strnlen(reinterpret_cast<const char*>(data), (size_t)len));
            }
        } else {
            throw H5Error(path + ": not a dataset of strings");
        }
        return out;
    }

    bool has_attribute(const std::string& obj, const std::string& name) const {
        uint64_t a;
        if (!locate(obj, a)) return false;
        Attr at;
        return find_attr(a, name, at);
    }
    /// numeric attribute (scalar or array) as doubles
    std::vector<double> attribute(const std::string& obj, const std::string& name) const {
        uint64_t a;
        if (!locate(obj, a)) throw H5Error("no such object: " + obj);
        Attr at;
        if (!find_attr(a, name, at)) throw H5Error("no attribute " + name + " on " + obj);
        std::vector<double> out(at.d.count());
        convert(at.data, at.d, out.data(), out.size(), obj + "@" + name);
        return out;
    }
    std::vector<std::string> attribute_names(const std::string& obj) const {
        uint64_t a;
        if (!locate(obj, a)) throw H5Error("no such object: " + obj);
        std::vector<std::string> out;
        for (auto& m : messages(a))
            if (m.type == 0x000C) {
                Attr at;
                std::string nm;
                parse_attr(m.p, m.n, nm, at);
                out.push_back(nm);
            }
        return out;
    }

  private:
    std::vector<uint8_t> buf_;
    uint64_t base_ = 0, root_header_ = 0;
    int so_ = 8, sl_ = 8;

    struct Msg { unsigned type; const uint8_t* p; size_t n; };
    struct DsRaw {
        Dataset d;
        int layout = -1;  // 0 compact, 1 contiguous, 2 chunked
        const uint8_t* compact = nullptr;
        uint64_t addr = 0, size = 0;
        std::vector<uint32_t> chunk;  // chunk dimensions (elements), rank entries
        uint64_t btree = 0;
        std::vector<int> filters;  // filter ids in pipeline order
    };
    struct Attr { Dataset d; const uint8_t* data = nullptr; };

    static const uint64_t UNDEF = ~(uint64_t)0;

    const uint8_t* at(uint64_t addr, size_t need) const {
        const uint64_t size = buf_.size();
        if (addr == UNDEF || base_ > size || addr > size - base_) throw H5Error("address out of file");
        const uint64_t a = base_ + addr;
        if (need > size - a) throw H5Error("address out of file");
        return buf_.data() + a;
    }
    /// a damaged file must not walk B-trees forever (cycles): nodes visited per read
    mutable uint64_t visits_ = 0;
    void visit() const {
        if (++visits_ > 2000000) throw H5Error("B-tree too large or cyclic");
    }
    /// p .. p + k must lie inside the message [m, m + n)
    static void need(const uint8_t* p, size_t k, const uint8_t* m, size_t n, const std::string& what) {
        if (p < m || (size_t)(p - m) > n || k > n - (size_t)(p - m)) throw H5Error(what + ": message too short");
    }
    static uint64_t rd(const uint8_t* p, int n) {
        uint64_t v = 0;
        for (int i = n - 1; i >= 0; --i) v = (v << 8) | p[i];
        return v;
    }
    uint64_t rdo(const uint8_t*& p) const { const uint64_t v = (so_ == 8 && rd(p, 8) == UNDEF) ? UNDEF : rd(p, so_); p += so_; return v; }
    uint64_t rdl(const uint8_t*& p) const { const uint64_t v = rd(p, sl_); p += sl_; return v; }

    void parse_superblock() {
        static const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        uint64_t off = 0;
        for (;; off = off ? off * 2 : 512) {  // the superblock may sit at 0, 512, 1024, ...
            if (off + 8 > buf_.size()) throw H5Error("not an HDF5 file (no signature)");
            if (!std::memcmp(buf_.data() + off, sig, 8)) break;
        }
        if (off + 128 > buf_.size()) throw H5Error("truncated file (superblock)");  // (the largest superblock has 96 bytes)
        const uint8_t* p = buf_.data() + off + 8;
        const int ver = p[0];
        if (ver == 0 || ver == 1) {
            so_ = p[5];
            sl_ = p[6];
            if ((so_ != 8 && so_ != 4) || (sl_ != 8 && sl_ != 4)) throw H5Error("unsupported size of offsets / lengths");
            p += 8 + 2 + 2 + 4;  // versions, sizes, group K values, flags
            if (ver == 1) p += 4;
            base_ = rdo(p);
            rdo(p);  // free-space info
            rdo(p);  // end of file
            rdo(p);  // driver info
            p += so_;  // root symbol table entry: link name offset
            root_header_ = rdo(p);
        } else if (ver == 2 || ver == 3) {
            so_ = p[1];
            sl_ = p[2];
            p += 4;
            base_ = rdo(p);
            rdo(p);  // superblock extension
            rdo(p);  // end of file
            root_header_ = rdo(p);
        } else {
            throw H5Error("unsupported superblock version " + std::to_string(ver));
        }
        if (base_ == 0) base_ = off;  // (addresses are relative to the base address)
    }

    /// all messages of a version-1 object header, continuation blocks followed
    std::vector<Msg> messages(uint64_t addr) const {
        const uint8_t* h = at(addr, 16);
        if (!std::memcmp(h, "OHDR", 4))
            throw H5Error("version-2 object header (file written with libver='latest'): not supported");
        if (h[0] != 1) throw H5Error("unsupported object header version " + std::to_string(h[0]));
        const unsigned nmsg = (unsigned)rd(h + 2, 2);
        const uint64_t size = rd(h + 8, 4);
        std::vector<Msg> out;
        std::vector<std::pair<uint64_t, uint64_t>> blocks;
        blocks.emplace_back(addr + 16, size);
        for (size_t b = 0; b < blocks.size() && out.size() < nmsg; ++b) {
            const uint8_t* p = at(blocks[b].first, (size_t)blocks[b].second);
            const uint8_t* end = p + blocks[b].second;
            while (p + 8 <= end && out.size() < nmsg) {
                const unsigned type = (unsigned)rd(p, 2);
                const size_t sz = (size_t)rd(p + 2, 2);
                const uint8_t* body = p + 8;
                if (body + sz > end) throw H5Error("object header message overruns its block");
                if (type == 0x0010) {  // continuation
                    const uint8_t* q = body;
                    const uint64_t coff = rdo(q), clen = rdl(q);
                    blocks.emplace_back(coff, clen);
                }
                out.push_back(Msg{type, body, sz});
                p = body + sz;
            }
        }
        return out;
    }

    /// object `index` of the global heap collection at `addr` (GCOL: version 1; objects = index u16, reference count
    /// u16, reserved u32, size, data padded to 8 bytes; object 0 is the free space)
    const uint8_t* global_heap_object(uint64_t addr, uint64_t index, size_t& nbytes, const std::string& what) const {
        const uint8_t* h = at(addr, 8 + sl_);
        if (std::memcmp(h, "GCOL", 4) || h[4] != 1) throw H5Error(what + ": bad global heap collection");
        const uint64_t total = rd(h + 8, sl_);
        h = at(addr, total);
        uint64_t off = 8 + sl_;
        while (off + 8 + sl_ <= total) {
            const uint64_t idx = rd(h + off, 2);
            const uint64_t size = rd(h + off + 8, sl_);
            if (idx == 0) break;  // free space: end of the objects
            if (off + 8 + sl_ + size > total) throw H5Error(what + ": global heap object overruns its collection");
            if (idx == index) {
                nbytes = (size_t)size;
                return h + off + 8 + sl_;
            }
            off += 8 + sl_ + ((size + 7) & ~(uint64_t)7);
        }
        throw H5Error(what + ": global heap object not found");
    }

    std::string heap_string(uint64_t heap_addr, uint64_t off) const {
        const uint8_t* h = at(heap_addr, 8 + 2 * sl_ + so_);
        if (std::memcmp(h, "HEAP", 4)) throw H5Error("bad local heap signature");
        const uint8_t* p = h + 8;
        const uint64_t seg_size = rdl(p);
        rdl(p);
        const uint64_t seg = rdo(p);
        if (off >= seg_size) throw H5Error("link name outside the local heap");
        const char* s = reinterpret_cast<const char*>(at(seg + off, 1));
        const size_t maxlen = (size_t)(seg_size - off);
        return std::string(s, strnlen(s, maxlen));
    }

    void btree_group(uint64_t node, uint64_t heap, std::map<std::string, uint64_t>& out, int depth = 0) const {
        if (depth > 32) throw H5Error("group B-tree too deep");
        visit();
        const uint8_t* h = at(node, 8 + 2 * so_);
        if (std::memcmp(h, "TREE", 4) || h[4] != 0) throw H5Error("bad group B-tree node");
        const int level = h[5];
        const unsigned used = (unsigned)rd(h + 6, 2);
        const uint8_t* p = h + 8 + 2 * so_;
        at(node, (size_t)(8 + 2 * so_ + (2 * used + 1) * (size_t)sl_ + used * (size_t)so_));  // bounds
        p += sl_;  // key 0
        for (unsigned i = 0; i < used; ++i) {
            const uint64_t child = rdo(p);
            p += sl_;  // key i + 1
            if (level > 0) {
                btree_group(child, heap, out, depth + 1);
            } else {
                const uint8_t* s = at(child, 8);
                if (std::memcmp(s, "SNOD", 4)) throw H5Error("bad symbol table node");
                const unsigned nsym = (unsigned)rd(s + 6, 2);
                const size_t esz = 2 * (size_t)so_ + 4 + 4 + 16;
                const uint8_t* e = at(child + 8, nsym * esz);
                for (unsigned k = 0; k < nsym; ++k, e += esz) {
                    const uint8_t* q = e;
                    const uint64_t name_off = rdo(q);
                    const uint64_t obj = rdo(q);
                    out[heap_string(heap, name_off)] = obj;
                }
            }
        }
    }

    std::map<std::string, uint64_t> children(uint64_t addr) const {
        std::map<std::string, uint64_t> out;
        for (auto& m : messages(addr)) {
            if (m.type == 0x0011) {
                const uint8_t* p = m.p;
                const uint64_t bt = rdo(p), heap = rdo(p);
                btree_group(bt, heap, out);
                return out;
            }
            if (m.type == 0x0002 || m.type == 0x0006)
                throw H5Error("new-style group (link messages): not supported, write the file with the default libver");
        }
        throw H5Error("object is not a group");
    }

    bool locate(const std::string& path, uint64_t& addr) const {
        addr = root_header_;
        size_t i = 0;
        while (i < path.size()) {
            while (i < path.size() && path[i] == '/') ++i;
            size_t j = i;
            while (j < path.size() && path[j] != '/') ++j;
            if (j == i) break;
            const std::string name = path.substr(i, j - i);
            std::map<std::string, uint64_t> ch;
            try {
                ch = children(addr);
            } catch (H5Error const&) {
                return false;
            }
            auto it = ch.find(name);
            if (it == ch.end()) return false;
            addr = it->second;
            i = j;
        }
        return true;
    }

    void parse_dataspace(const uint8_t* p, size_t n, Dataset& d, const std::string& what) const {
        if (n < 4) throw H5Error(what + ": short dataspace message");
        const int ver = p[0];
        d.rank = p[1];
        const uint8_t* q;
        if (ver == 1) q = p + 8;
        else if (ver == 2) {
            if (p[3] == 2) throw H5Error(what + ": null dataspace");
            q = p + 4;
        } else throw H5Error(what + ": unsupported dataspace version");
        d.dims.clear();
        if (d.rank > 32) throw H5Error(what + ": dataspace rank");
        need(q, (size_t)d.rank * sl_, p, n, what);
        for (int i = 0; i < d.rank; ++i) d.dims.push_back(rdl(q));
    }
    void parse_datatype(const uint8_t* p, size_t n, Dataset& d, const std::string& what) const {
        if (n < 8) throw H5Error(what + ": short datatype message");
        d.type_class = p[0] & 0x0f;
        d.type_size = (int)rd(p + 4, 4);
        if (d.type_class == 0) {
            if (p[1] & 1) throw H5Error(what + ": big-endian integers are not supported");
            d.type_signed = (p[1] & 8) != 0;
            if (d.type_size != 1 && d.type_size != 2 && d.type_size != 4 && d.type_size != 8) throw H5Error(what + ": integer size");
        } else if (d.type_class == 1) {
            if (p[1] & 1) throw H5Error(what + ": big-endian floats are not supported");
            if (d.type_size != 4 && d.type_size != 8) throw H5Error(what + ": float size");
        } else if (d.type_class == 3) {
            // fixed-length string: readable as raw bytes only
        } else if (d.type_class == 9) {
            // variable-length: strings (bits 0-3 of the class bit field = 1) or sequences of integers
            if ((p[1] & 0x0f) == 1) {
                d.vlen_string = true;
                return;
            }
            if ((p[1] & 0x0f) != 0) throw H5Error(what + ": unknown variable-length type");
            if (n < 16) throw H5Error(what + ": short variable-length datatype message");
            Dataset base;
            parse_datatype(p + 8, n - 8, base, what + " (element type)");
            if (base.type_class != 0) throw H5Error(what + ": variable-length sequences of non-integer elements are not supported");
            d.base_class = base.type_class;
            d.base_size = base.type_size;
            d.base_signed = base.type_signed;
        } else {
            throw H5Error(what + ": datatype class " + std::to_string(d.type_class) + " is not supported");
        }
    }

    DsRaw dataset_raw(uint64_t addr, const std::string& what) const {
        DsRaw r;
        bool have_space = false, have_type = false;
        for (auto& m : messages(addr)) {
            if (m.type == 0x0001) { parse_dataspace(m.p, m.n, r.d, what); have_space = true; }
            else if (m.type == 0x0003) { parse_datatype(m.p, m.n, r.d, what); have_type = true; }
            else if (m.type == 0x0008) {
                const uint8_t* p = m.p;
                need(p, 2, m.p, m.n, what);
                if (p[0] != 3) throw H5Error(what + ": data layout message version " + std::to_string(p[0]) + " is not supported");
                r.layout = p[1];
                p += 2;
                if (r.layout == 0) {
                    need(p, 2, m.p, m.n, what);
                    r.size = rd(p, 2);
                    r.compact = p + 2;
                    need(r.compact, (size_t)r.size, m.p, m.n, what);
                } else if (r.layout == 1) {
                    need(p, (size_t)so_ + sl_, m.p, m.n, what);
                    r.addr = rdo(p);
                    r.size = rdl(p);
                } else if (r.layout == 2) {
                    need(p, 1 + (size_t)so_, m.p, m.n, what);
                    const int dim = *p++;
                    r.btree = rdo(p);
                    if (dim < 1 || dim > 33) throw H5Error(what + ": chunk rank");
                    need(p, 4 * (size_t)(dim - 1), m.p, m.n, what);
                    for (int i = 0; i < dim - 1; ++i, p += 4) r.chunk.push_back((uint32_t)rd(p, 4));
                } else throw H5Error(what + ": unknown layout class");
            } else if (m.type == 0x000B) {
                const uint8_t* p = m.p;
                need(p, 2, m.p, m.n, what);
                const int ver = p[0], nf = p[1];
                p += (ver == 1) ? 8 : 2;
                for (int i = 0; i < nf; ++i) {
                    need(p, 8, m.p, m.n, what);
                    const int id = (int)rd(p, 2);
                    p += 2;
                    int name_len = 0;
                    if (ver == 1 || id >= 256) { name_len = (int)rd(p, 2); p += 2; }
                    p += 2;  // flags
                    const int ncd = (int)rd(p, 2);
                    p += 2;
                    if (ver == 1) name_len = (name_len + 7) & ~7;
                    p += name_len;
                    p += 4 * ncd;
                    if (ver == 1 && (ncd & 1)) p += 4;
                    r.filters.push_back(id);
                }
            }
        }
        if (!have_space || !have_type || r.layout < 0) throw H5Error(what + ": not a dataset");
        return r;
    }

    // ---- zlib, resolved at run time (only files with compressed chunks need it)
    typedef int (*uncompress_fn)(unsigned char*, unsigned long*, const unsigned char*, unsigned long);
    static uncompress_fn zlib_uncompress() {
        static uncompress_fn fn = []() -> uncompress_fn {
            void* h = dlopen("libz.so.1", RTLD_NOW);
            if (!h) h = dlopen("libz.so", RTLD_NOW);
            return h ? reinterpret_cast<uncompress_fn>(dlsym(h, "uncompress")) : nullptr;
        }();
        return fn;
    }

    void chunk_tree(uint64_t node, const DsRaw& r, std::vector<uint8_t>& out, const std::string& what, int depth = 0) const {
        if (depth > 32) throw H5Error(what + ": chunk B-tree too deep");
        visit();
        const uint8_t* h = at(node, 8 + 2 * so_);
        if (std::memcmp(h, "TREE", 4) || h[4] != 1) throw H5Error(what + ": bad chunk B-tree node");
        const int level = h[5];
        const unsigned used = (unsigned)rd(h + 6, 2);
        const int rank = r.d.rank;
        const size_t keysz = 8 + 8 * (size_t)(rank + 1);
        const uint8_t* p = at(node + 8 + 2 * so_, (used + 1) * keysz + used * (size_t)so_);
        const size_t es = (size_t)r.d.type_size;
        uint64_t chunk_elems64 = 1, chunk_bytes64 = 0;
        for (int i = 0; i < rank; ++i)
            if (r.chunk[i] == 0 || __builtin_mul_overflow(chunk_elems64, (uint64_t)r.chunk[i], &chunk_elems64))
                throw H5Error(what + ": chunk dimensions");
        // (a chunk is at most as large as the -- already bounded -- dataset, rounded up to whole chunks per dimension:
        // bounded here by the output size times 2^rank, and absolutely)
        if (__builtin_mul_overflow(chunk_elems64, (uint64_t)es, &chunk_bytes64) || chunk_bytes64 > ((uint64_t)1 << 31) ||
            chunk_bytes64 > (uint64_t)buf_.size() * 1100u + (16u << 20))
            throw H5Error(what + ": chunk larger than its file can hold");
        const size_t chunk_elems = (size_t)chunk_elems64;
        for (unsigned i = 0; i < used; ++i) {
            const uint32_t csize = (uint32_t)rd(p, 4);
            const uint32_t fmask = (uint32_t)rd(p + 4, 4);
            std::vector<uint64_t> coff(rank);
            for (int k = 0; k < rank; ++k) coff[k] = rd(p + 8 + 8 * k, 8);
            p += keysz;
            const uint64_t child = rdo(p);
            if (level > 0) {
                chunk_tree(child, r, out, what, depth + 1);
                continue;
            }
            // ---- one chunk: undo the filters (last applied first), then scatter into the output
            std::vector<uint8_t> data(at(child, csize), at(child, csize) + csize);
            for (int f = (int)r.filters.size() - 1; f >= 0; --f) {
                if (fmask & (1u << f)) continue;  // filter skipped for this chunk
                const int id = r.filters[f];
                if (id == 1) {
                    uncompress_fn un = zlib_uncompress();
                    if (!un) throw H5Error(what + ": compressed chunks need libz.so.1, which could not be loaded");
                    std::vector<uint8_t> raw(chunk_elems * es);
                    unsigned long dl = (unsigned long)raw.size();
                    if (un(raw.data(), &dl, data.data(), (unsigned long)data.size()) != 0) throw H5Error(what + ": inflate failed");
                    raw.resize(dl);
                    data.swap(raw);
                } else if (id == 2) {
                    const size_t ne = data.size() / es;
                    std::vector<uint8_t> un(data.size());
                    for (size_t b = 0; b < es; ++b)
                        for (size_t e = 0; e < ne; ++e) un[e * es + b] = data[b * ne + e];
                    for (size_t e = ne * es; e < data.size(); ++e) un[e] = data[e];
                    data.swap(un);
                } else if (id == 3) {
                    if (data.size() >= 4) data.resize(data.size() - 4);  // fletcher32 checksum (not verified)
                } else {
                    throw H5Error(what + ": filter " + std::to_string(id) + " is not supported");
                }
            }
            if (data.size() < chunk_elems * es) throw H5Error(what + ": chunk shorter than its dimensions");
            // copy the part of the chunk that lies inside the dataset (row-major, last dimension fastest)
            std::vector<uint64_t> idx(rank, 0);
            for (;;) {
                bool inside = true;
                uint64_t dst = 0, src = 0;
                for (int k = 0; k < rank; ++k) {
                    const uint64_t g = coff[k] + idx[k];
                    if (g >= r.d.dims[k]) inside = false;
                    dst = dst * r.d.dims[k] + g;
                    src = src * r.chunk[k] + idx[k];
                }
                if (inside) std::memcpy(out.data() + dst * es, data.data() + src * es, es);
                int k = rank - 1;
                for (; k >= 0; --k) {
                    if (++idx[k] < r.chunk[k]) break;
                    idx[k] = 0;
                }
                if (k < 0) break;
            }
        }
    }

    std::vector<uint8_t> read_raw(const DsRaw& r, const std::string& what) const {
        // size of the dataset in memory, bounded by what the file can hold (a damaged dataspace or datatype must end
        // in an error, not in a huge allocation): contiguous / compact data are in the file as they are, chunks may be
        // deflated (ratio bounded generously)
        uint64_t bytes64 = 0;
        if (r.d.type_size <= 0 || __builtin_mul_overflow(r.d.count(), (uint64_t)r.d.type_size, &bytes64))
            throw H5Error(what + ": dataset size overflows");
        const uint64_t limit = r.layout == 2 ? (uint64_t)buf_.size() * 1100u + (16u << 20) : (uint64_t)buf_.size();
        if (bytes64 > limit) throw H5Error(what + ": dataset larger than its file can hold");
        const size_t bytes = (size_t)bytes64;
        std::vector<uint8_t> out(bytes, 0);
        if (bytes == 0) return out;
        if (r.layout == 0) {
            if (r.size < bytes) throw H5Error(what + ": compact data shorter than the dataspace");
            std::memcpy(out.data(), r.compact, bytes);
        } else if (r.layout == 1) {
            if (r.addr == UNDEF) return out;  // never written: fill value 0
            std::memcpy(out.data(), at(r.addr, bytes), bytes);
        } else {
            if ((int)r.chunk.size() != r.d.rank) throw H5Error(what + ": chunk rank mismatch");
            if (r.btree != UNDEF) chunk_tree(r.btree, r, out, what);
        }
        return out;
    }

    static int64_t get_int(const uint8_t* p, int size, bool sgn) {
        uint64_t v = rd(p, size);
        if (sgn && size < 8 && (v >> (8 * size - 1))) v |= ~(uint64_t)0 << (8 * size);
        return (int64_t)v;
    }
    static void convert(const uint8_t* raw, const Dataset& d, double* out, size_t n, const std::string& what) {
        for (size_t i = 0; i < n; ++i) {
            const uint8_t* p = raw + i * d.type_size;
            if (d.type_class == 1) {
                if (d.type_size == 8) { double v; std::memcpy(&v, p, 8); out[i] = v; }
                else { float v; std::memcpy(&v, p, 4); out[i] = v; }
            } else if (d.type_class == 0) {
                const int64_t v = get_int(p, d.type_size, d.type_signed);
                out[i] = d.type_signed ? (double)v : (double)(uint64_t)v;
            } else {
                throw H5Error(what + ": not a numeric type");
            }
        }
    }

    void parse_attr(const uint8_t* p, size_t n, std::string& name, Attr& a) const {
        need(p, 9, p, n, "attribute");
        const int ver = p[0];
        if (ver < 1 || ver > 3) throw H5Error("unsupported attribute message version");
        const size_t name_sz = (size_t)rd(p + 2, 2), dt_sz = (size_t)rd(p + 4, 2), ds_sz = (size_t)rd(p + 6, 2);
        const uint8_t* q = p + 8;
        if (ver == 3) ++q;  // name character set
        auto pad = [&](size_t v) { return ver == 1 ? ((v + 7) & ~(size_t)7) : v; };
        need(q, pad(name_sz), p, n, "attribute");
        name.assign(reinterpret_cast<const char*>(q), strnlen(reinterpret_cast<const char*>(q), name_sz));
        q += pad(name_sz);
        need(q, pad(dt_sz), p, n, "attribute " + name);
        parse_datatype(q, dt_sz, a.d, "attribute " + name);
        q += pad(dt_sz);
        need(q, pad(ds_sz), p, n, "attribute " + name);
        if (ds_sz >= 8 && q[0] == 1 && q[1] == 0) { a.d.rank = 0; a.d.dims.clear(); }  // scalar (v1)
        else parse_dataspace(q, ds_sz, a.d, "attribute " + name);
        q += pad(ds_sz);
        a.data = q;
        uint64_t nbytes = 0;
        if (a.d.type_size <= 0 || __builtin_mul_overflow(a.d.count(), (uint64_t)a.d.type_size, &nbytes))
            throw H5Error("attribute " + name + ": size overflows");
        need(q, (size_t)nbytes, p, n, "attribute " + name + ": data");
    }
    bool find_attr(uint64_t addr, const std::string& name, Attr& a) const {
        for (auto& m : messages(addr))
            if (m.type == 0x000C) {
                std::string nm;
                Attr t;
                parse_attr(m.p, m.n, nm, t);
                if (nm == name) { a = t; return true; }
            }
        return false;
    }
};

}  // namespace pb200
#endif
