// streambench.cu -- what HBM bandwidth does a B200 deliver for the fused step's traffic SHAPE?
// 13 read streams + 8 write streams of doubles as a flat grid-stride sweep and as the fused kernel's
// pattern (CTAs own column strips of 8240-double rows and march north through a y-chunk, 4 layers x 21 planes),
// with variations of the strip width, request shape (per-warp pieces / whole segments / TMA bulk) and read:write mix.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/streambench tools/streambench.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

struct Ptrs { const double *r[13]; double *w[8]; };

template <int NR, int NW>
__global__ void k_flat(Ptrs P, size_t n2) {  // n2 = number of double2 elements per stream
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += (size_t)gridDim.x * blockDim.x) {
    double2 acc = make_double2(0.0, 0.0);
#pragma unroll
    for (int s = 0; s < NR; s++) {
      const double2 v = __ldg(reinterpret_cast<const double2 *>(P.r[s]) + i);
      acc.x += v.x; acc.y += v.y;
    }
#pragma unroll
    for (int s = 0; s < NW; s++) __stcs(reinterpret_cast<double2 *>(P.w[s]) + i, make_double2(acc.x + s, acc.y));
  }
}

// strips: CTA (bx, by) owns columns [bx*4*USE, +4*USE) of rows [by*rows_per, ...) for all 4 layers; warp = (layer, USE-column group)
template <int NR, int NW, int USE, int ST = 0>
__global__ void k_strips(Ptrs P, int NX, int NY, size_t plane, int rows_per, int xoff = 0) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, grp = wid & 3, l = wid >> 2;
  const int x = xoff + blockIdx.x * 4 * USE + grp * USE + lane;
  const bool on = lane < USE && x < NX;
  const int y0 = blockIdx.y * rows_per, y1 = min(y0 + rows_per, NY);
  const size_t L = (size_t)l * plane;
  for (int y = y0; y < y1; y++) {
    const size_t c = L + (size_t)y * NX + x;
    double acc = 0.0;
    if (on) {
#pragma unroll
      for (int s = 0; s < NR; s++) acc += __ldg(P.r[s] + c);
#pragma unroll
      for (int s = 0; s < NW; s++) {
        if (ST == 0) __stcs(P.w[s] + c, acc + s);
        else if (ST == 1) P.w[s][c] = acc + s;
        else if (ST == 2) __stcg(P.w[s] + c, acc + s);
        else if (ST == 3) __stwt(P.w[s] + c, acc + s);
      }
    }
  }
}

// store-shape variants of the 4 x 28-column strips (loads as in k_strips):
//  VAR 0: 14 lanes store 16 bytes each (same 224 bytes per warp)      VAR 1: warp g writes columns [32g, 32g+28) (no line shared by two warps)
//  VAR 2: the four 28-column results go through shared memory and ONE warp per layer writes the 112 columns with full 32-lane stores
template <int NR, int NW, int VAR>
__global__ void k_strips_var(Ptrs P, int NX, int NY, size_t plane, int rows_per) {
  __shared__ double stage[4][NW][112];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5, grp = wid & 3, l = wid >> 2;
  const int cw = (VAR == 1) ? 32 : 28;
  const int x = 16 + blockIdx.x * 4 * cw + grp * cw + lane;
  const bool on = lane < 28 && x < NX;
  const int y0 = blockIdx.y * rows_per, y1 = min(y0 + rows_per, NY);
  const size_t L = (size_t)l * plane;
  for (int y = y0; y < y1; y++) {
    const size_t c = L + (size_t)y * NX + x;
    double acc = 0.0;
    if (on) {
#pragma unroll
      for (int s = 0; s < NR; s++) acc += __ldg(P.r[s] + c);
    }
    if (VAR == 0) {
      const double nb = __shfl_down_sync(0xffffffffu, acc, 1);
      if (on && !(lane & 1)) {
#pragma unroll
        for (int s = 0; s < NW; s++) __stcs(reinterpret_cast<double2 *>(P.w[s] + c), make_double2(acc + s, nb + s));
      }
    } else if (VAR == 1) {
      if (on) {
#pragma unroll
        for (int s = 0; s < NW; s++) __stcs(P.w[s] + c, acc + s);
      }
    } else {
      if (lane < 28) {
#pragma unroll
        for (int s = 0; s < NW; s++) stage[l][s][grp * 28 + lane] = acc + s;
      }
      __syncthreads();
      if (grp == 0) {
        const size_t c0 = L + (size_t)y * NX + 16 + blockIdx.x * 112;
#pragma unroll
        for (int s = 0; s < NW; s++)
          for (int k = lane; k < 112; k += 32)
            if (16 + blockIdx.x * 112 + k < NX) __stcs(P.w[s] + c0 + k, stage[l][s][k]);
      }
      __syncthreads();
    }
  }
}

// wide: one warp per layer moves the CTA's whole 128-column row segment (1 KiB) with 32-byte lanes
template <int NR, int NW>
__global__ void k_wide(Ptrs P, int NX, int NY, size_t plane, int rows_per) {
  const int lane = threadIdx.x & 31, l = threadIdx.x >> 5;
  const int x = blockIdx.x * 128 + lane * 4;
  const bool on = x + 3 < NX;
  const int y0 = blockIdx.y * rows_per, y1 = min(y0 + rows_per, NY);
  const size_t L = (size_t)l * plane;
  for (int y = y0; y < y1; y++) {
    const size_t c = L + (size_t)y * NX + x;
    double4 acc = make_double4(0, 0, 0, 0);
    if (on) {
#pragma unroll
      for (int s = 0; s < NR; s++) {
        const double2 a = __ldg(reinterpret_cast<const double2 *>(P.r[s] + c)), b = __ldg(reinterpret_cast<const double2 *>(P.r[s] + c + 2));
        acc.x += a.x; acc.y += a.y; acc.z += b.x; acc.w += b.y;
      }
#pragma unroll
      for (int s = 0; s < NW; s++) {
        __stcs(reinterpret_cast<double2 *>(P.w[s] + c), make_double2(acc.x + s, acc.y));
        __stcs(reinterpret_cast<double2 *>(P.w[s] + c + 2), make_double2(acc.z, acc.w));
      }
    }
  }
}

// TMA: per layer one warp; lane s stages stream s's 960-byte row segment (bulk global->shared, 2 rows in flight),
// then lanes < NW write segments back with bulk shared->global copies
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
template <int NR, int NW, int DEPTH>
__global__ void k_tma(Ptrs P, int NX, int NY, size_t plane, int rows_per, int segcols) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, l = threadIdx.x >> 5, nl = blockDim.x >> 5;
  const unsigned segb = segcols * 8;
  unsigned char *ring = smem + (size_t)l * DEPTH * NR * segb;              // [DEPTH][NR][segb]
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)nl * DEPTH * NR * segb) + l * DEPTH;
  const int x = blockIdx.x * segcols;
  if (x + segcols > NX) return;
  const int y0 = blockIdx.y * rows_per, y1 = min(y0 + rows_per, NY);
  const size_t L = (size_t)l * plane;
  if (lane == 0) for (int k = 0; k < DEPTH; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + k)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  auto issue = [&](int y) {
    const int k = (y - y0) % DEPTH;
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + k)), "r"(NR * segb) : "memory");
    __syncwarp();
    if (lane < NR)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + ((size_t)k * NR + lane) * segb)),
                   "l"(P.r[lane] + L + (size_t)y * NX + x), "r"(segb), "r"(smem_u32(bars + k)) : "memory");
  };
  for (int d = 0; d < DEPTH - 1 && y0 + d < y1; d++) issue(y0 + d);
  for (int y = y0; y < y1; y++) {
    const int k = (y - y0) % DEPTH;
    const unsigned par = ((y - y0) / DEPTH) & 1;
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bars + k)), "r"(par) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane < NW) {
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(P.w[lane] + L + (size_t)y * NX + x), "r"(smem_u32(ring + ((size_t)k * NR + lane) * segb)), "r"(segb) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    if (y + DEPTH - 1 < y1) issue(y + DEPTH - 1);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// the fused step's proposed shape: stage 120-column segments starting 4 columns left of the CTA's 112 result columns
// (x0 = 16 + 112*bx, 128-byte aligned), write the 112 result columns back from the same slots, either as one
// 896-byte bulk store or as four 224-byte pieces (PIECES = 4)
template <int NR, int NW, int DEPTH, int PIECES>
__global__ void k_tma_fused(Ptrs P, int NX, int NY, size_t plane, int rows_per) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int lane = threadIdx.x & 31, l = threadIdx.x >> 5, nl = blockDim.x >> 5;
  const unsigned segb = 120 * 8;
  unsigned char *ring = smem + (size_t)l * DEPTH * NR * segb;
  unsigned long long *bars = reinterpret_cast<unsigned long long *>(smem + (size_t)nl * DEPTH * NR * segb) + l * DEPTH;
  const int x0 = 16 + blockIdx.x * 112;
  if (x0 + 116 > NX) return;
  const int y0 = blockIdx.y * rows_per, y1 = min(y0 + rows_per, NY);
  const size_t L = (size_t)l * plane;
  if (lane == 0) for (int k = 0; k < DEPTH; k++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(bars + k)));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncwarp();
  auto issue = [&](int y) {
    const int k = (y - y0) % DEPTH;
    if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bars + k)), "r"(NR * segb) : "memory");
    __syncwarp();
    if (lane < NR)
      asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(ring + ((size_t)k * NR + lane) * segb)),
                   "l"(P.r[lane] + L + (size_t)y * NX + x0 - 4), "r"(segb), "r"(smem_u32(bars + k)) : "memory");
  };
  for (int d = 0; d < DEPTH - 1 && y0 + d < y1; d++) issue(y0 + d);
  for (int y = y0; y < y1; y++) {
    const int k = (y - y0) % DEPTH;
    const unsigned par = ((y - y0) / DEPTH) & 1;
    asm volatile("{\n\t.reg .pred p;\n\tW_%=:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D_%=;\n\tbra W_%=;\n\tD_%=:\n\t}" ::"r"(smem_u32(bars + k)), "r"(par) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    if (lane < NW * PIECES) {
      const int s = lane / PIECES, pc = lane % PIECES;
      const unsigned pb = 896 / PIECES;
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(P.w[s] + L + (size_t)y * NX + x0 + pc * (112 / PIECES)),
                   "r"(smem_u32(ring + ((size_t)k * NR + s) * segb + 32 + pc * pb)), "r"(pb) : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
    if (y + DEPTH - 1 < y1) issue(y + DEPTH - 1);
  }
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main(int argc, char **argv) {
  const int NX = 8240, NY = 8201, NL = 4;
  const size_t plane = (size_t)NX * NY, n = plane * NL;
  Ptrs P;
  for (int s = 0; s < 13; s++) { double *p; CK(cudaMalloc(&p, n * 8)); CK(cudaMemset(p, 0, n * 8)); P.r[s] = p; }
  for (int s = 0; s < 8; s++) { CK(cudaMalloc(&P.w[s], n * 8)); CK(cudaMemset(P.w[s], 0, n * 8)); }
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  auto timeit = [&](const char *name, double bytes, auto launch) {
    for (int i = 0; i < 2; i++) launch();
    CK(cudaDeviceSynchronize());
    float best = 1e30f, tot = 0;
    for (int i = 0; i < 6; i++) {
      CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = ms < best ? ms : best; tot += ms;
    }
    CK(cudaGetLastError());
    printf("%-52s best %7.3f ms %7.1f GB/s   mean %7.1f GB/s\n", name, best, bytes / best / 1e6, bytes / (tot / 6) / 1e6);
    fflush(stdout);
  };
  const double b1 = (double)n * 8;
  const int rp4 = (NY + 3) / 4, rp16 = (NY + 15) / 16;
  timeit("flat 13R+8W", 21 * b1, [&] { k_flat<13, 8><<<148 * 16, 512>>>(P, n / 2); });
  timeit("strips 4x28 cols 13R+8W, x0 = 4 (fused kernel today)", 21 * b1 * 8192 / 8240, [&] { k_strips<13, 8, 28><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4, 4); });
  timeit("strips 4x28 cols 13R+8W, x0 = 16 (CTA 128B-aligned)", 21 * b1 * 8176 / 8240, [&] { k_strips<13, 8, 28><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4, 16); });
  timeit("strips 4x28 13R+8W, 14 lanes x 16 B stores", 21 * b1 * 8176 / 8240, [&] { k_strips_var<13, 8, 0><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4); });
  timeit("strips 4x28 13R+8W, warps 32 columns apart", 21 * b1 * 7168 / 8240, [&] { k_strips_var<13, 8, 1><<<dim3(64, 4), 512>>>(P, NX, NY, plane, rp4); });
  timeit("strips 4x28 13R+8W, smem + one warp stores 112 cols", 21 * b1 * 8176 / 8240, [&] { k_strips_var<13, 8, 2><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4); });
  timeit("strips 4x28 13R+8W plain st.global", 21 * b1 * 8176 / 8240, [&] { k_strips<13, 8, 28, 1><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4, 4); });
  timeit("strips 4x28 13R+8W st.global.cg", 21 * b1 * 8176 / 8240, [&] { k_strips<13, 8, 28, 2><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4, 4); });
  timeit("strips 4x28 13R+8W st.global.wt", 21 * b1 * 8176 / 8240, [&] { k_strips<13, 8, 28, 3><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4, 4); });
  timeit("strips 4x28 13R+8W plain st, 16 chunks", 21 * b1 * 8176 / 8240, [&] { k_strips<13, 8, 28, 1><<<dim3(73, 16), 512>>>(P, NX, NY, plane, rp16, 4); });
  timeit("strips 4x28 1R+8W plain st", 9 * b1 * 8176 / 8240, [&] { k_strips<1, 8, 28, 1><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4, 4); });
  timeit("strips 4x28 cols 1R+8W,  x0 = 16", 9 * b1 * 8176 / 8240, [&] { k_strips<1, 8, 28><<<dim3(73, 4), 512>>>(P, NX, NY, plane, rp4, 16); });
  {
    const size_t sh = (size_t)4 * 2 * 13 * 960 + 4 * 2 * 8;
    const double by = 4.0 * (double)NY * 73 * (13 * 960.0 + 8 * 896.0);
    CK(cudaFuncSetAttribute(k_tma_fused<13, 8, 2, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    CK(cudaFuncSetAttribute(k_tma_fused<13, 8, 2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
    timeit("TMA fused shape: 120-col loads, 896 B stores, depth 2", by, [&] { k_tma_fused<13, 8, 2, 1><<<dim3(73, 4), 128, sh>>>(P, NX, NY, plane, rp4); });
    timeit("TMA fused shape: 120-col loads, 4x224 B stores, depth 2", by, [&] { k_tma_fused<13, 8, 2, 4><<<dim3(73, 4), 128, sh>>>(P, NX, NY, plane, rp4); });
    timeit("TMA fused shape: 896 B stores, depth 2, 8 chunks", by, [&] { k_tma_fused<13, 8, 2, 1><<<dim3(73, 8), 128, sh>>>(P, NX, NY, plane, (NY + 7) / 8); });
  }
  return 0;
}
