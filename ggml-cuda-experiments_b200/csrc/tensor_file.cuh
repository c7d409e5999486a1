// tensor_file.cuh — host-only reader/writer of the ggml tensor-dump files the reference replays as fixtures
// (fa-cuda-{q,k,v,mask,qkv}-256.tensor; loader: reference src/utils.h:110-150, use: src/flash-matrix.cu:69-73).
//
// File layout, little endian, no padding:
//     i32 n_dims | i32 type (0 = f32, 1 = f16) | i32 ne[n_dims] (fastest dimension first) | i32 name_len | name bytes
//     (no terminator) | raw data, ne[0] fastest
// The reference's loader keeps the name in a char[20] and never returns ne; this reader returns both and refuses what
// that loader could not hold (name_len > 19) only on the WRITE side, so files written here always load there.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/b200fa.h"

namespace b200fa {

struct TfFile {
    FILE* f = nullptr;
    ~TfFile() { if (f) fclose(f); }
};

inline bool tf_read_i32(FILE* f, int32_t& v) { return fread(&v, 1, sizeof(v), f) == sizeof(v); }

inline int tf_parse_header(FILE* f, b200fa_tensor_info* info) {
    memset(info, 0, sizeof(*info));
    for (int i = 0; i < 4; i++) info->ne[i] = 1;
    int32_t n_dims, type, len;
    if (!tf_read_i32(f, n_dims) || !tf_read_i32(f, type)) return B200FA_ERR_IO;
    if (n_dims < 1 || n_dims > 4) return B200FA_ERR_INVALID;
    if (type != B200FA_TYPE_F32 && type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    int64_t n = 1;
    for (int i = 0; i < n_dims; i++) {
        int32_t e;
        if (!tf_read_i32(f, e)) return B200FA_ERR_IO;
        if (e < 0) return B200FA_ERR_INVALID;
        info->ne[i] = e;
        n *= e;
        if (n > ((int64_t)1 << 40)) return B200FA_ERR_INVALID;
    }
    if (!tf_read_i32(f, len)) return B200FA_ERR_IO;
    if (len < 0 || len >= (int32_t)sizeof(info->name)) return B200FA_ERR_INVALID;
    if (len > 0 && fread(info->name, 1, (size_t)len, f) != (size_t)len) return B200FA_ERR_IO;
    info->name[len] = '\0';
    info->n_dims = n_dims;
    info->type = type;
    info->data_offset = (int64_t)sizeof(int32_t) * (3 + n_dims) + len;
    info->data_bytes = n * (type == B200FA_TYPE_F16 ? 2 : 4);
    return B200FA_OK;
}

inline int tf_info(const char* path, b200fa_tensor_info* info) {
    if (path == nullptr || info == nullptr) return B200FA_ERR_INVALID;
    TfFile h;
    h.f = fopen(path, "rb");
    if (!h.f) return B200FA_ERR_IO;
    const int rc = tf_parse_header(h.f, info);
    if (rc != B200FA_OK) return rc;
    // the payload must be there in full: a truncated capture is an error, not a short read
    if (fseek(h.f, 0, SEEK_END) != 0) return B200FA_ERR_IO;
    const long end = ftell(h.f);
    if (end < 0 || (int64_t)end < info->data_offset + info->data_bytes) return B200FA_ERR_IO;
    return B200FA_OK;
}

inline int tf_read(const char* path, void* dst, size_t dst_bytes) {
    if (path == nullptr || dst == nullptr) return B200FA_ERR_INVALID;
    TfFile h;
    h.f = fopen(path, "rb");
    if (!h.f) return B200FA_ERR_IO;
    b200fa_tensor_info info;
    const int rc = tf_parse_header(h.f, &info);
    if (rc != B200FA_OK) return rc;
    if ((int64_t)dst_bytes < info.data_bytes) return B200FA_ERR_INVALID;
    if (fread(dst, 1, (size_t)info.data_bytes, h.f) != (size_t)info.data_bytes) return B200FA_ERR_IO;
    return B200FA_OK;
}

inline int tf_write(const char* path, const char* name, int type, int n_dims, const int64_t* ne, const void* data) {
    if (path == nullptr || name == nullptr || ne == nullptr) return B200FA_ERR_INVALID;
    if (n_dims < 1 || n_dims > 4) return B200FA_ERR_INVALID;
    if (type != B200FA_TYPE_F32 && type != B200FA_TYPE_F16) return B200FA_ERR_UNSUPPORTED;
    const size_t len = strlen(name);
    if (len > 19) return B200FA_ERR_INVALID;  // the reference loader's name field is char[20] (utils.h:107)
    int64_t n = 1;
    for (int i = 0; i < n_dims; i++) {
        if (ne[i] < 0 || ne[i] > INT32_MAX) return B200FA_ERR_INVALID;
        n *= ne[i];
    }
    if (n > 0 && data == nullptr) return B200FA_ERR_INVALID;
    TfFile h;
    h.f = fopen(path, "wb");
    if (!h.f) return B200FA_ERR_IO;
    int32_t hdr[8];
    int k = 0;
    hdr[k++] = n_dims;
    hdr[k++] = type;
    for (int i = 0; i < n_dims; i++) hdr[k++] = (int32_t)ne[i];
    hdr[k++] = (int32_t)len;
    const size_t bytes = (size_t)n * (type == B200FA_TYPE_F16 ? 2 : 4);
    if (fwrite(hdr, sizeof(int32_t), (size_t)k, h.f) != (size_t)k) return B200FA_ERR_IO;
    if (len > 0 && fwrite(name, 1, len, h.f) != len) return B200FA_ERR_IO;
    if (bytes > 0 && fwrite(data, 1, bytes, h.f) != bytes) return B200FA_ERR_IO;
    FILE* f = h.f;
    h.f = nullptr;
    return fclose(f) == 0 ? B200FA_OK : B200FA_ERR_IO;
}

}  // namespace b200fa
