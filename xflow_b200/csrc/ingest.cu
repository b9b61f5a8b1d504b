// Device side of the ingest: feature-id hashing on the GPU.
//
// The reference's loader turns the id token of "field:id:val" into the parameter-server key with
// std::hash<std::string> (load_data_from_disk.cc:151,173,194).  For numeric ids — what the bundled data
// and every synthetic config use — the key is the MurmurHash64A (hash.h) of the id's decimal string.
// A caller that already holds integer ids can therefore ship 4 bytes per token instead of 8-byte keys
// and let the device make the keys: halves the PCIe traffic of the end-to-end path, bit-exact by
// construction (same hash.h function, __host__ __device__).
#include <cuda_runtime.h>
#include <stdint.h>

#include "hash.h"
#include "internal.h"

__global__ void xf_k_hash_ids(const uint32_t* __restrict__ ids, uint32_t n, uint64_t* __restrict__ keys) {
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    uint32_t v = __ldcs(ids + i);
    char rev[12];
    int len = 0;
    do { rev[len++] = (char)('0' + v % 10u); v /= 10u; } while (v);
    char buf[12];
    for (int j = 0; j < len; ++j) buf[j] = rev[len - 1 - j];
    keys[i] = xf_murmur64a(buf, (uint64_t)len);
  }
}

int xf_launch_hash_ids(const uint32_t* d_ids, uint32_t n, uint64_t* d_keys, cudaStream_t st) {
  if (n == 0) return XF_OK;
  xf_k_hash_ids<<<xf_grid_for(n, 256, 8), 256, 0, st>>>(d_ids, n, d_keys);
  XF_CUDA_TRY(cudaGetLastError());
  return XF_OK;
}

XF_DLL int xf_hash_decimal_ids_device(const uint32_t* d_ids, uint64_t n, uint64_t* d_keys, void* cuda_stream) {
  if ((!d_ids || !d_keys) && n) return XF_ERR_ARG;
  if (n > 0xFFFFFFFFull) return XF_ERR_ARG;
  return xf_launch_hash_ids(d_ids, (uint32_t)n, d_keys, (cudaStream_t)cuda_stream);
}
