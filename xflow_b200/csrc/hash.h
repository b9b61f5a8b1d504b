// Feature-id hashing of the reference's loader: std::hash<std::string> (io.h:46, used at
// load_data_from_disk.cc:151,173,194).  With libstdc++ on a 64-bit target that is _Hash_bytes, the
// MurmurHash64A variant with seed 0xc70f6907 (libstdc++-v3/libsupc++/hash_bytes.cc, GCC 13).
// The keys it produces ARE the parameter-server keys, so this must be bit-exact; the device copy in
// ingest.cu uses the same function through __host__ __device__.
#pragma once
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define XF_HD __host__ __device__ __forceinline__
#else
#define XF_HD inline
#endif

XF_HD uint64_t xf_murmur_load(const char* p, int nbytes) {
  // little-endian load of the first nbytes (1..8) bytes, unaligned-safe
  uint64_t v = 0;
  for (int i = nbytes - 1; i >= 0; --i) v = (v << 8) | (uint64_t)(unsigned char)p[i];
  return v;
}

XF_HD uint64_t xf_murmur64a(const char* s, uint64_t len) {
  const uint64_t m = 0xc6a4a7935bd1e995ull;
  const int r = 47;
  uint64_t h = 0xc70f6907ull ^ (len * m);
  const uint64_t nfull = len & ~(uint64_t)7;
  for (uint64_t off = 0; off < nfull; off += 8) {
    uint64_t k = xf_murmur_load(s + off, 8);
    k *= m;
    k ^= k >> r;
    k *= m;
    h ^= k;
    h *= m;
  }
  const int tail = (int)(len & 7);
  if (tail) {
    h ^= xf_murmur_load(s + nfull, tail);
    h *= m;
  }
  h ^= h >> r;
  h *= m;
  h ^= h >> r;
  return h;
}
