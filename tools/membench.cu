// Micro-benchmarks behind the round-2 kernel decisions: which access PRIMITIVES limit the table kernels?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/membench tools/membench.cu
//   tools/membench [table_MB=8192] [accesses_M=6.5]
// Table of 32-byte rows (the LR row), `n` distinct-ish random rows per launch (one batch's worth).
//   read        one 256-bit load per access
//   rmw         load + 256-bit store
//   lazy        load + CAS on the tag word + 256-bit store + f64 RED   (the "open + accumulate" of step_lazy.cu)
//   lazy_sync   the same with __syncwarp between the stages (warp-synchronous, like the row kernel)
//   red         f64 RED only
//   pf+read     prefetch.global.L2 of n2 rows in one kernel, then the read kernel over the same rows
//   chain-k     k dependent loads per access (probe chains)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

__device__ __forceinline__ uint64_t mix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__device__ __forceinline__ void ld256(const uint8_t* p, uint64_t& a, uint64_t& b, uint64_t& c, uint64_t& d) {
  asm volatile("ld.global.cg.v4.u64 {%0,%1,%2,%3}, [%4];" : "=l"(a), "=l"(b), "=l"(c), "=l"(d) : "l"(p));
}
__device__ __forceinline__ void st256(uint8_t* p, uint64_t a, uint64_t b, uint64_t c, uint64_t d) {
  asm volatile("st.global.v4.u64 [%0], {%1,%2,%3,%4};" ::"l"(p), "l"(a), "l"(b), "l"(c), "l"(d) : "memory");
}

enum { M_READ = 0, M_RMW, M_LAZY, M_LAZY_SYNC, M_RED, M_PF, M_CHAIN2, M_CHAIN4, M_READ64, M_READ128, M_LD2, M_ST, M_LD_RED, M_RED2, M_CAS, M_LD_CAS_ST, M_LD_ST_RED, M_ST16, M_RED_F32, M_LD_ST16, M_CAS128, M_LD_CAS128, M_LD_CAS128_RED };

template <int MODE>
__global__ void __launch_bounds__(256) k(uint8_t* base, uint64_t mask, uint64_t n, uint64_t seed, uint32_t tag, uint64_t* sink) {
  uint64_t acc = 0;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
    uint8_t* p = base + ((mix(seed + i) & mask) << 5);
    uint64_t a, b, c, d;
    if (MODE == M_PF) {
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    } else if (MODE == M_RED) {
      atomicAdd(reinterpret_cast<double*>(p + 24), 1.0);
    } else if (MODE == M_READ) {
      ld256(p, a, b, c, d);
      acc += a ^ b ^ c ^ d;
    } else if (MODE == M_READ64 || MODE == M_READ128) {
      // the whole 64 / 128-byte aligned group around the row (bucket probing)
      const int nb = MODE == M_READ64 ? 2 : 4;
      uint8_t* q = (uint8_t*)((uint64_t)p & ~(uint64_t)(nb * 32 - 1));
      for (int j = 0; j < nb; ++j) { ld256(q + 32 * j, a, b, c, d); acc += a ^ b ^ c ^ d; }
    } else if (MODE == M_LD2) {          // the same sector twice, the second load depends on the first
      ld256(p, a, b, c, d);
      uint8_t* q = p + ((a >> 63) << 5);  // a is 0 or small: q == p, but the address depends on the data
      ld256(q, a, b, c, d);
      acc += a ^ d;
    } else if (MODE == M_ST) {           // blind full-sector store (no load)
      st256(p, i, 1, 2, 3);
    } else if (MODE == M_ST16) {         // blind 16-byte store (half a sector)
      *reinterpret_cast<uint4*>(p + 16) = make_uint4(1, 2, 3, (uint32_t)i);
    } else if (MODE == M_LD_ST16) {      // load + 16-byte store
      ld256(p, a, b, c, d);
      *reinterpret_cast<uint4*>(p + 16) = make_uint4((uint32_t)a, 2, 3, (uint32_t)i);
    } else if (MODE == M_LD_RED) {
      ld256(p, a, b, c, d);
      atomicAdd(reinterpret_cast<double*>(p + 24), 1.0 + (double)(a & 1));
    } else if (MODE == M_RED2) {
      atomicAdd(reinterpret_cast<double*>(p + 24), 1.0);
      atomicAdd(reinterpret_cast<double*>(p + 24), 2.0);
    } else if (MODE == M_RED_F32) {
      atomicAdd(reinterpret_cast<float*>(p + 24), 1.0f);
    } else if (MODE == M_CAS) {
      acc += atomicCAS(reinterpret_cast<unsigned int*>(p + 20), 0u, 7u);
    } else if (MODE == M_LD_CAS_ST) {
      ld256(p, a, b, c, d);
      const uint32_t old_tag = (uint32_t)(c >> 32);
      const uint32_t got = atomicCAS(reinterpret_cast<unsigned int*>(p + 20), old_tag, 0xFFFFFFFFu);
      if (got == old_tag) st256(p, a, b + 1, (c & 0xFFFFFFFFull) | ((uint64_t)tag << 32), 0);
    } else if (MODE == M_LD_ST_RED) {
      ld256(p, a, b, c, d);
      st256(p, a, b + 1, (c & 0xFFFFFFFFull) | ((uint64_t)tag << 32), 0);
      atomicAdd(reinterpret_cast<double*>(p + 24), 1.0);
    } else if (MODE == M_CAS128 || MODE == M_LD_CAS128 || MODE == M_LD_CAS128_RED) {
      // claim + publish in ONE 128-bit compare-and-swap on the row's second half {w,n,z,tag}
      c = 0; d = 0;
      if (MODE != M_CAS128) ld256(p, a, b, c, d);
      uint64_t o0, o1;
      const uint64_t n0 = c + 1, n1 = (d & 0xFFFFFFFFull) | ((uint64_t)tag << 32);
      asm volatile("{\n .reg .b128 cmp, swp, old;\n mov.b128 cmp, {%2, %3};\n mov.b128 swp, {%4, %5};\n"
                   " atom.global.cas.b128 old, [%6], cmp, swp;\n mov.b128 {%0, %1}, old;\n}"
                   : "=l"(o0), "=l"(o1) : "l"(c), "l"(d), "l"(n0), "l"(n1), "l"(p + 16) : "memory");
      acc += o0 ^ o1;
      if (MODE == M_LD_CAS128_RED) atomicAdd(reinterpret_cast<unsigned long long*>(p + 8), 12345ull + (o0 & 1));
    } else if (MODE == M_RMW) {
      ld256(p, a, b, c, d);
      st256(p, a + 1, b, c, d);
    } else if (MODE == M_LAZY || MODE == M_LAZY_SYNC) {
      ld256(p, a, b, c, d);
      const uint32_t old_tag = (uint32_t)(c >> 32);
      if (MODE == M_LAZY_SYNC) __syncwarp();
      const uint32_t got = atomicCAS(reinterpret_cast<unsigned int*>(p + 20), old_tag, 0xFFFFFFFFu);
      if (MODE == M_LAZY_SYNC) __syncwarp();
      if (got == old_tag) st256(p, a, b + 1, (c & 0xFFFFFFFFull) | ((uint64_t)tag << 32), 0);
      if (MODE == M_LAZY_SYNC) __syncwarp();
      atomicAdd(reinterpret_cast<double*>(p + 24), 1.0);
    } else if (MODE == M_CHAIN2 || MODE == M_CHAIN4) {
      const int kk = MODE == M_CHAIN2 ? 2 : 4;
      for (int j = 0; j < kk; ++j) {
        ld256(p, a, b, c, d);
        acc += a;
        p = base + ((mix(seed + i + (a & 1) + 1000003ull * (j + 1)) & mask) << 5);  // depends on the loaded value
      }
    }
  }
  if (acc == 0x123456789ull) *sink = acc;
}

// V: 0 read 4 lanes x 4 x uint4 | 1 read 16 lanes x uint4 | 2 rmw 4 lanes | 3 rmw 16 lanes
template <int V>
__global__ void __launch_bounds__(256) krow(uint8_t* base, uint64_t mask, uint64_t n, uint64_t seed, uint64_t* sink) {
  const int G = (V & 1) ? 16 : 4;
  const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;
  const uint64_t ngrp = (uint64_t)gridDim.x * blockDim.x / G;
  const int q = threadIdx.x % G;
  uint32_t acc = 0;
  for (uint64_t i = gid; i < n; i += ngrp) {
    uint4* row = reinterpret_cast<uint4*>(base + ((mix(seed + i) & mask) << 8));
    if (G == 16) {
      uint4 v = __ldcg(row + q);
      acc += v.x ^ v.w;
      if (V >= 2) { v.x += 1; row[q] = v; }
    } else {
      uint4 v[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) v[j] = __ldcg(row + 4 * j + q);
#pragma unroll
      for (int j = 0; j < 4; ++j) { acc += v[j].x ^ v[j].w; if (V >= 2) { v[j].x += 1; row[4 * j + q] = v[j]; } }
    }
  }
  if (acc == 0x12345u) *sink = acc;
}

template <int MODE>
static float run(const char* name, uint8_t* base, uint64_t nsect, uint64_t n, uint64_t* sink, int bps = 8, uint64_t seed0 = 77,
                 bool print = true, int reps = 3) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    k<MODE><<<148 * bps, 256>>>(base, nsect - 1, n, seed0 + 1000 * r, 5 + r, sink);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  if (print) printf("  %-40s CTAs/SM %d  %8.1f us   %7.2f G accesses/s\n", name, bps, best * 1e3, (double)n / (best * 1e-3) / 1e9);
  return best;
}

int main(int argc, char** argv) {
  uint64_t mb = argc > 1 ? strtoull(argv[1], 0, 10) : 8192;
  double acc_m = argc > 2 ? atof(argv[2]) : 6.5;
  uint64_t n = (uint64_t)(acc_m * 1e6);
  uint64_t nsect = 1;
  while (nsect * 32 < mb * 1048576ull) nsect <<= 1;
  uint8_t* base;
  uint64_t* sink;
  cudaMalloc(&base, nsect * 32);
  cudaMalloc(&sink, 8);
  cudaMemset(base, 0, nsect * 32);
  for (int fetch = 32; fetch <= 128; fetch *= 2) {
    if (fetch == 64) continue;
    cudaDeviceSetLimit(cudaLimitMaxL2FetchGranularity, fetch);
    printf("table %llu MiB, %llu accesses per launch, L2 fetch granularity %d B\n", (unsigned long long)(nsect * 32 >> 20),
           (unsigned long long)n, fetch);
    for (int bps = 2; bps <= 8; bps *= 2) run<M_READ>("read", base, nsect, n, sink, bps);
    run<M_READ64>("read 64-B group (2 loads)", base, nsect, n, sink);
    run<M_READ128>("read 128-B group (4 loads)", base, nsect, n, sink);
    for (int bps = 4; bps <= 8; bps *= 2) run<M_RMW>("rmw (load + store)", base, nsect, n, sink, bps);
    run<M_RED>("red f64", base, nsect, n, sink);
    for (int bps = 4; bps <= 8; bps *= 2) run<M_LAZY>("lazy (load + CAS + store + RED)", base, nsect, n, sink, bps);
    run<M_LAZY_SYNC>("lazy, warp-synchronous stages", base, nsect, n, sink);
    run<M_LD2>("load the same row twice (dependent)", base, nsect, n, sink);
    run<M_ST>("blind 32-B store", base, nsect, n, sink);
    run<M_ST16>("blind 16-B store", base, nsect, n, sink);
    run<M_LD_ST16>("load + 16-B store", base, nsect, n, sink);
    run<M_LD_RED>("load + RED", base, nsect, n, sink);
    run<M_RED2>("RED twice on the row", base, nsect, n, sink);
    run<M_RED_F32>("red f32", base, nsect, n, sink);
    run<M_CAS>("CAS only (returning)", base, nsect, n, sink);
    run<M_LD_CAS_ST>("load + CAS + store", base, nsect, n, sink);
    run<M_LD_ST_RED>("load + store + RED", base, nsect, n, sink);
    run<M_CAS128>("CAS.128 only", base, nsect, n, sink);
    run<M_LD_CAS128>("load + CAS.128", base, nsect, n, sink);
    run<M_LD_CAS128_RED>("load + CAS.128 + RED.u64", base, nsect, n, sink);
    run<M_CHAIN2>("chain of 2 dependent loads", base, nsect, 2 * n / 2, sink);
    run<M_CHAIN4>("chain of 4 dependent loads", base, nsect, n, sink);
    // prefetch effectiveness: rows that fit L2 (64 MB), prefetched by one kernel and read by the next
    {
      const uint64_t n2 = 2000000;
      float cold = 1e30f, warm = 1e30f, pf = 1e30f;
      for (int r = 0; r < 3; ++r) {
        run<M_RED>("", base, nsect, 8000000, sink, 8, 9000 + r, false, 1);  // wipe L2 with other rows
        float c = run<M_READ>("", base, nsect, n2, sink, 8, 555 + r, false, 1);
        run<M_RED>("", base, nsect, 8000000, sink, 8, 9100 + r, false, 1);
        float p = run<M_PF>("", base, nsect, n2, sink, 8, 777 + r, false, 1);
        float w = run<M_READ>("", base, nsect, n2, sink, 8, 777 + r, false, 1);
        if (c < cold) cold = c;
        if (w < warm) warm = w;
        if (p < pf) pf = p;
      }
      printf("  prefetch.global.L2 check (2 M rows = 64 MB): cold read %.1f us, prefetch kernel %.1f us, read after prefetch %.1f us\n",
             cold * 1e3, pf * 1e3, warm * 1e3);
    }
  }
  // 256-byte rows (the FM k=16 FTRL row): per-thread pieces vs one cooperative instruction per row
  {
    const uint64_t nrow = nsect / 8;
    printf("256-B rows, %llu rows, %llu random rows per launch\n", (unsigned long long)nrow, (unsigned long long)n);
    void (*kern[4])(uint8_t*, uint64_t, uint64_t, uint64_t, uint64_t*) = {krow<0>, krow<1>, krow<2>, krow<3>};
    const char* names[4] = {"read row: 4 lanes x 4 x 16 B (as xf_k_update)", "read row: 16 lanes x 16 B, one instruction",
                            "rmw row: 4 lanes x 4 x 16 B", "rmw row: 16 lanes x 16 B, one instruction"};
    for (int v = 0; v < 4; ++v) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0); cudaEventCreate(&e1);
      float best = 1e30f;
      for (int r = 0; r < 3; ++r) {
        cudaEventRecord(e0);
        kern[v]<<<148 * 8, 256>>>(base, nrow - 1, n, 99 + r, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
      }
      printf("  %-52s %8.1f us   %7.2f G rows/s\n", names[v], best * 1e3, (double)n / (best * 1e-3) / 1e9);
    }
  }
  return 0;
}
