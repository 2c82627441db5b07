// h5wmin.h -- minimal HDF5 WRITER without an HDF5 library: what the result logging of the hot path needs
// (the counterpart of the reader h5min.h).
//
// Writes the classic file format every HDF5 library reads: superblock version 0 (8-byte offsets and lengths),
// version-1 object headers, groups as symbol tables (one leaf B-tree node + one symbol node per group + a local
// heap), contiguous datasets of IEEE doubles, 32/64-bit little-endian integers and fixed-length strings,
// numeric attributes.  No chunking, no filters, no extendable datasets: the caller collects its records in
// memory and writes the file once (the reference appends to extendable datasets through the HDF5 library,
// src/parpecommon/hdf5Misc.cpp hdf5CreateOrExtendAndWriteTo*; the resulting datasets have the same names,
// types and shapes).
#ifndef PB200_H5WMIN_H
#define PB200_H5WMIN_H

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace pb200 {

class H5Writer {
  public:
    enum Type { F64, I32, I64, STR };
    struct Dataset {
        Type type = F64;
        std::vector<uint64_t> dims;
        std::vector<uint8_t> bytes;
        int str_len = 0;  // STR: bytes per element (null padded)
    };
    struct Node {
        std::map<std::string, std::unique_ptr<Node>> groups;
        std::map<std::string, Dataset> datasets;
        std::map<std::string, Dataset> attrs;
    };

    Node& group(const std::string& path) {
        Node* n = &root_;
        for (auto& part : split(path)) {
            auto& slot = n->groups[part];
            if (!slot) slot = std::make_unique<Node>();
            n = slot.get();
        }
        return *n;
    }
    void doubles(const std::string& path, const std::vector<double>& v, std::vector<uint64_t> dims) {
        Dataset d;
        d.type = F64;
        d.dims = std::move(dims);
        d.bytes.resize(v.size() * 8);
        if (!v.empty()) std::memcpy(d.bytes.data(), v.data(), d.bytes.size());
        check(d, v.size(), path);
        put(path, std::move(d));
    }
    void ints(const std::string& path, const std::vector<int32_t>& v, std::vector<uint64_t> dims) {
        Dataset d;
        d.type = I32;
        d.dims = std::move(dims);
        d.bytes.resize(v.size() * 4);
        if (!v.empty()) std::memcpy(d.bytes.data(), v.data(), d.bytes.size());
        check(d, v.size(), path);
        put(path, std::move(d));
    }
    void int64s(const std::string& path, const std::vector<int64_t>& v, std::vector<uint64_t> dims) {
        Dataset d;
        d.type = I64;
        d.dims = std::move(dims);
        d.bytes.resize(v.size() * 8);
        if (!v.empty()) std::memcpy(d.bytes.data(), v.data(), d.bytes.size());
        check(d, v.size(), path);
        put(path, std::move(d));
    }
    /// 1-D dataset of fixed-length, null-padded strings
    void strings(const std::string& path, const std::vector<std::string>& v) {
        Dataset d;
        d.type = STR;
        size_t len = 1;
        for (auto& s : v) len = s.size() + 1 > len ? s.size() + 1 : len;
        d.str_len = (int)len;
        d.dims = {v.size()};
        d.bytes.assign(v.size() * len, 0);
        for (size_t i = 0; i < v.size(); ++i) std::memcpy(d.bytes.data() + i * len, v[i].data(), v[i].size());
        put(path, std::move(d));
    }
    void attr_double(const std::string& group_path, const std::string& name, double v) {
        Dataset d;
        d.type = F64;
        d.dims = {1};
        d.bytes.resize(8);
        std::memcpy(d.bytes.data(), &v, 8);
        group(group_path).attrs[name] = std::move(d);
    }

    void save(const std::string& path) {
        buf_.assign(96, 0);  // superblock (56 bytes) + root symbol table entry (40 bytes)
        // one symbol node per group: the node size 2 * leaf_k must hold the largest group
        size_t maxc = max_children(root_);
        uint64_t leaf_k = (maxc + 1) / 2 + 1;
        if (leaf_k < 4) leaf_k = 4;
        if (leaf_k >= 65536) throw std::runtime_error("hdf5 writer: too many children in one group");
        leaf_k_ = leaf_k;
        uint64_t tree = 0, heap = 0;
        const uint64_t root = write_group(root_, tree, heap);
        std::vector<uint8_t> sb;
        const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
        sb.insert(sb.end(), sig, sig + 8);
        const uint8_t ver[8] = {0, 0, 0, 0, 0, 8, 8, 0};
        sb.insert(sb.end(), ver, ver + 8);
        u16(sb, (uint16_t)leaf_k); u16(sb, 16); u32(sb, 0);
        u64(sb, 0); u64(sb, UNDEF); u64(sb, buf_.size()); u64(sb, UNDEF);
        u64(sb, 0); u64(sb, root); u32(sb, 1); u32(sb, 0); u64(sb, tree); u64(sb, heap);
        if (sb.size() != 96) throw std::runtime_error("hdf5 writer: superblock size");
        std::memcpy(buf_.data(), sb.data(), 96);
        FILE* f = std::fopen(path.c_str(), "wb");
        if (!f) throw std::runtime_error("hdf5 writer: cannot open " + path);
        const size_t n = std::fwrite(buf_.data(), 1, buf_.size(), f);
        std::fclose(f);
        if (n != buf_.size()) throw std::runtime_error("hdf5 writer: short write to " + path);
    }

  private:
    static constexpr uint64_t UNDEF = ~(uint64_t)0;
    Node root_;
    std::vector<uint8_t> buf_;
    uint64_t leaf_k_ = 4;

    static std::vector<std::string> split(const std::string& path) {
        std::vector<std::string> parts;
        std::string cur;
        for (char c : path) {
            if (c == '/') {
                if (!cur.empty()) parts.push_back(cur);
                cur.clear();
            } else cur.push_back(c);
        }
        if (!cur.empty()) parts.push_back(cur);
        return parts;
    }
    static void check(const Dataset& d, size_t n, const std::string& path) {
        uint64_t c = 1;
        for (auto v : d.dims) c *= v;
        if (c != n) throw std::runtime_error("hdf5 writer: " + path + ": dimensions do not match the data");
    }
    void put(const std::string& path, Dataset d) {
        auto parts = split(path);
        if (parts.empty()) throw std::runtime_error("hdf5 writer: empty dataset path");
        std::string g;
        for (size_t i = 0; i + 1 < parts.size(); ++i) g += "/" + parts[i];
        group(g).datasets[parts.back()] = std::move(d);
    }
    static void u16(std::vector<uint8_t>& b, uint16_t v) { for (int i = 0; i < 2; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    static void u32(std::vector<uint8_t>& b, uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    static void u64(std::vector<uint8_t>& b, uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
    static void pad8(std::vector<uint8_t>& b) { while (b.size() % 8) b.push_back(0); }

    uint64_t alloc(const std::vector<uint8_t>& data) {
        while (buf_.size() % 8) buf_.push_back(0);
        const uint64_t addr = buf_.size();
        buf_.insert(buf_.end(), data.begin(), data.end());
        return addr;
    }
    static std::vector<uint8_t> datatype(const Dataset& d) {
        std::vector<uint8_t> t;
        if (d.type == F64) {
            // class 1 (floating point) version 1; little endian, implied mantissa normalisation, sign bit 63
            const uint8_t h[4] = {0x11, 0x20, 63, 0};
            t.insert(t.end(), h, h + 4);
            u32(t, 8);
            u16(t, 0); u16(t, 64); t.push_back(52); t.push_back(11); t.push_back(0); t.push_back(52); u32(t, 1023);
        } else if (d.type == I32 || d.type == I64) {
            const uint32_t sz = d.type == I32 ? 4 : 8;
            const uint8_t h[4] = {0x10, 0x08, 0, 0};  // class 0 (fixed point), little endian, signed
            t.insert(t.end(), h, h + 4);
            u32(t, sz);
            u16(t, 0); u16(t, (uint16_t)(8 * sz));
        } else {
            const uint8_t h[4] = {0x13, 0x00, 0, 0};  // class 3 (string): null terminated, ASCII
            t.insert(t.end(), h, h + 4);
            u32(t, (uint32_t)d.str_len);
        }
        return t;
    }
    static std::vector<uint8_t> dataspace(const std::vector<uint64_t>& dims) {
        std::vector<uint8_t> s = {1, (uint8_t)dims.size(), 0, 0, 0, 0, 0, 0};
        for (auto v : dims) u64(s, v);
        return s;
    }
    static std::vector<uint8_t> message(uint16_t type, std::vector<uint8_t> body) {
        pad8(body);
        std::vector<uint8_t> m;
        u16(m, type); u16(m, (uint16_t)body.size());
        m.push_back(0); m.push_back(0); m.push_back(0); m.push_back(0);
        m.insert(m.end(), body.begin(), body.end());
        return m;
    }
    static std::vector<uint8_t> attr_message(const std::string& name, const Dataset& d) {
        std::vector<uint8_t> nm(name.begin(), name.end());
        nm.push_back(0);
        std::vector<uint8_t> dt = datatype(d), ds = dataspace(d.dims);
        std::vector<uint8_t> body = {1, 0};
        u16(body, (uint16_t)nm.size()); u16(body, (uint16_t)dt.size()); u16(body, (uint16_t)ds.size());
        pad8(nm); pad8(dt); pad8(ds);
        body.insert(body.end(), nm.begin(), nm.end());
        body.insert(body.end(), dt.begin(), dt.end());
        body.insert(body.end(), ds.begin(), ds.end());
        body.insert(body.end(), d.bytes.begin(), d.bytes.end());
        return message(0x000C, body);
    }
    uint64_t object_header(const std::vector<std::vector<uint8_t>>& msgs) {
        std::vector<uint8_t> body;
        for (auto& m : msgs) body.insert(body.end(), m.begin(), m.end());
        std::vector<uint8_t> h = {1, 0};
        u16(h, (uint16_t)msgs.size()); u32(h, 1); u32(h, (uint32_t)body.size());
        h.push_back(0); h.push_back(0); h.push_back(0); h.push_back(0);
        h.insert(h.end(), body.begin(), body.end());
        return alloc(h);
    }
    uint64_t write_dataset(const Dataset& d) {
        std::vector<uint64_t> dims = d.dims.empty() ? std::vector<uint64_t>{1} : d.dims;
        std::vector<std::vector<uint8_t>> msgs;
        msgs.push_back(message(0x0001, dataspace(dims)));
        msgs.push_back(message(0x0003, datatype(d)));
        const uint64_t addr = d.bytes.empty() ? UNDEF : alloc(d.bytes);
        std::vector<uint8_t> lay = {3, 1};  // layout message version 3, contiguous
        u64(lay, addr); u64(lay, d.bytes.size());
        msgs.push_back(message(0x0008, lay));
        return object_header(msgs);
    }
    uint64_t write_group(const Node& g, uint64_t& tree_addr, uint64_t& heap_addr) {
        // children in name order (the B-tree's key order), groups and datasets together
        std::map<std::string, uint64_t> entries;
        for (auto& kv : g.groups) {
            uint64_t t, h;
            entries[kv.first] = write_group(*kv.second, t, h);
        }
        for (auto& kv : g.datasets) entries[kv.first] = write_dataset(kv.second);
        // local heap: offset 0 holds the empty string (the B-tree's first key), then the names
        std::vector<uint8_t> heap(8, 0);
        std::vector<uint64_t> offs;
        for (auto& kv : entries) {
            offs.push_back(heap.size());
            heap.insert(heap.end(), kv.first.begin(), kv.first.end());
            heap.push_back(0);
            pad8(heap);
        }
        const uint64_t heap_data = alloc(heap);
        std::vector<uint8_t> hh = {'H', 'E', 'A', 'P', 0, 0, 0, 0};
        u64(hh, heap.size()); u64(hh, UNDEF); u64(hh, heap_data);
        heap_addr = alloc(hh);
        std::vector<uint8_t> snod = {'S', 'N', 'O', 'D', 1, 0};
        u16(snod, (uint16_t)entries.size());
        size_t i = 0;
        for (auto& kv : entries) {
            u64(snod, offs[i++]); u64(snod, kv.second); u32(snod, 0); u32(snod, 0);
            snod.insert(snod.end(), 16, 0);
        }
        snod.resize(8 + 2 * leaf_k_ * 40, 0);  // (nodes have their full on-disk size: 2 * leaf_k entries)
        const uint64_t snod_addr = alloc(snod);
        std::vector<uint8_t> tree = {'T', 'R', 'E', 'E', 0, 0};
        if (!entries.empty()) {
            u16(tree, 1); u64(tree, UNDEF); u64(tree, UNDEF);
            u64(tree, 0); u64(tree, snod_addr); u64(tree, offs.back());
        } else {
            u16(tree, 0); u64(tree, UNDEF); u64(tree, UNDEF); u64(tree, 0);
        }
        tree.resize(24 + (2 * 16 + 1) * 8 + 2 * 16 * 8, 0);  // (full node of the group B-tree: internal K = 16)
        tree_addr = alloc(tree);
        std::vector<std::vector<uint8_t>> msgs;
        std::vector<uint8_t> st;
        u64(st, tree_addr); u64(st, heap_addr);
        msgs.push_back(message(0x0011, st));
        for (auto& kv : g.attrs) msgs.push_back(attr_message(kv.first, kv.second));
        return object_header(msgs);
    }
    static size_t max_children(const Node& g) {
        size_t n = g.groups.size() + g.datasets.size();
        for (auto& kv : g.groups) {
            const size_t c = max_children(*kv.second);
            n = c > n ? c : n;
        }
        return n;
    }
};

}  // namespace pb200
#endif
