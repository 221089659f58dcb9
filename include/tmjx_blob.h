/*
 * tmjx_blob.h — reader for the flat model-constant table written by
 * track-mjx_b200/model_blob.py (the stand-in for `mjx.put_model(...)`, reference
 * track_mjx/environment/task/single_clip_tracking.py:91).
 *
 * Header-only, host-side C++; used by the CUDA library (csrc/) and by the CPU oracle (oracle/).
 *   header   : magic 'TMJX' (u32) | version (u32) | n_sections (i32) | reserved (i32)
 *   directory: n_sections x { name[24] | dtype (i32: 0=f32, 1=i32) | count (i32) | offset (i64) }
 */
#ifndef TMJX_BLOB_H_
#define TMJX_BLOB_H_

#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace tmjx {

constexpr uint32_t kBlobMagic = 0x584A4D54u;  // 'TMJX'
constexpr uint32_t kBlobVersion = 1u;

class Blob {
 public:
  Blob(const void* data, size_t n) : p_(static_cast<const uint8_t*>(data)), n_(n) {
    if (n < 16) throw std::runtime_error("model blob: truncated header");
    uint32_t magic, version;
    std::memcpy(&magic, p_, 4);
    std::memcpy(&version, p_ + 4, 4);
    std::memcpy(&nsec_, p_ + 8, 4);
    if (magic != kBlobMagic || version != kBlobVersion) throw std::runtime_error("model blob: bad magic/version");
    if (nsec_ < 0 || 16 + size_t(nsec_) * 40 > n) throw std::runtime_error("model blob: truncated directory");
  }

  bool has(const char* name) const { return find(name) >= 0; }

  std::vector<float> f32(const char* name) const {
    int i = need(name, 0);
    std::vector<float> v(count(i));
    if (!v.empty()) std::memcpy(v.data(), p_ + offset(i), v.size() * 4);
    return v;
  }
  std::vector<int32_t> i32(const char* name) const {
    int i = need(name, 1);
    std::vector<int32_t> v(count(i));
    if (!v.empty()) std::memcpy(v.data(), p_ + offset(i), v.size() * 4);
    return v;
  }

 private:
  const uint8_t* entry(int i) const { return p_ + 16 + size_t(i) * 40; }
  int dtype(int i) const { int32_t d; std::memcpy(&d, entry(i) + 24, 4); return d; }
  size_t count(int i) const { int32_t c; std::memcpy(&c, entry(i) + 28, 4); return size_t(c); }
  size_t offset(int i) const { int64_t o; std::memcpy(&o, entry(i) + 32, 8); return size_t(o); }
  int find(const char* name) const {
    for (int i = 0; i < nsec_; ++i)
      if (std::strncmp(reinterpret_cast<const char*>(entry(i)), name, 24) == 0) return i;
    return -1;
  }
  int need(const char* name, int dt) const {
    int i = find(name);
    if (i < 0) throw std::runtime_error(std::string("model blob: missing section ") + name);
    if (dtype(i) != dt) throw std::runtime_error(std::string("model blob: wrong dtype for ") + name);
    if (offset(i) + count(i) * 4 > n_) throw std::runtime_error(std::string("model blob: truncated section ") + name);
    return i;
  }
  const uint8_t* p_;
  size_t n_;
  int32_t nsec_;
};

}  // namespace tmjx
#endif  // TMJX_BLOB_H_
