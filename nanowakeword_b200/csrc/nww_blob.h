// nww_blob.h — parser for the packed weight blob handed across the C ABI (host C++).
//
// Layout (little endian), produced by nanowakeword_b200/weights.py::pack_blob:
//   char  magic[8] = "NWWB200\0";  u32 version = 1;  u32 n_tensors;
//   per tensor: u32 name_len; char name[name_len]; u32 dtype (0 = f32, 1 = i32); u32 ndim;
//               u32 dims[ndim]; u64 nbytes; zero padding up to a 16-byte boundary (from blob
//               start); nbytes of data.
#pragma once

#include <stdint.h>
#include <string.h>
#include <map>
#include <string>
#include <vector>

namespace nww {

struct BlobTensor {
    size_t offset = 0;          // byte offset of the data inside the blob
    size_t nbytes = 0;
    int dtype = 0;
    std::vector<uint32_t> dims;
    size_t numel() const {
        size_t n = 1;
        for (uint32_t d : dims) n *= d;
        return n;
    }
};

struct Blob {
    const unsigned char* base = nullptr;
    size_t size = 0;
    std::map<std::string, BlobTensor> tensors;

    bool has(const std::string& n) const { return tensors.count(n) != 0; }
    const BlobTensor* find(const std::string& n) const {
        auto it = tensors.find(n);
        return it == tensors.end() ? nullptr : &it->second;
    }
    const float* f32(const std::string& n) const {
        const BlobTensor* t = find(n);
        return t ? reinterpret_cast<const float*>(base + t->offset) : nullptr;
    }
};

inline bool parse_blob(const void* data, size_t size, Blob* out, std::string* err) {
    const unsigned char* p = static_cast<const unsigned char*>(data);
    auto fail = [&](const char* m) {
        if (err) *err = m;
        return false;
    };
    if (size < 16 || memcmp(p, "NWWB200", 8) != 0) return fail("weight blob: bad magic");
    uint32_t version, n;
    memcpy(&version, p + 8, 4);
    memcpy(&n, p + 12, 4);
    if (version != 1) return fail("weight blob: unsupported version");
    size_t off = 16;
    out->base = p;
    out->size = size;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t name_len;
        if (off + 4 > size) return fail("weight blob: truncated");
        memcpy(&name_len, p + off, 4);
        off += 4;
        if (off + name_len + 8 > size) return fail("weight blob: truncated");
        std::string name(reinterpret_cast<const char*>(p + off), name_len);
        off += name_len;
        BlobTensor t;
        uint32_t dtype, ndim;
        memcpy(&dtype, p + off, 4);
        memcpy(&ndim, p + off + 4, 4);
        off += 8;
        if (ndim > 8 || off + 4 * ndim + 8 > size) return fail("weight blob: bad tensor header");
        t.dtype = (int)dtype;
        t.dims.resize(ndim);
        memcpy(t.dims.data(), p + off, 4 * ndim);
        off += 4 * ndim;
        uint64_t nbytes;
        memcpy(&nbytes, p + off, 8);
        off += 8;
        off = (off + 15) & ~(size_t)15;
        if (off + nbytes > size) return fail("weight blob: tensor data out of range");
        if (nbytes != t.numel() * 4) return fail("weight blob: size/shape mismatch");
        t.offset = off;
        t.nbytes = (size_t)nbytes;
        off += nbytes;
        out->tensors[name] = t;
    }
    return true;
}

}  // namespace nww
