// Backward of the FeaSt aggregation (training step, train_dual.py:199-218).
//
// Forward (feast.cu):  P = X U^T;  l_ijh = P_jh - P_ih + c_h;  q = softmax_h(l);  Z_i[h,:] = 1/d_i * sum_{j in N(i)+{i}} q_ijh x_j
// Given dZ [N, 9*C]:
//   dq_ijh  = 1/d_i * <dZ_i[h,:], x_j>                      dl_ijh = q_ijh (dq_ijh - sum_h' q_ijh' dq_ijh')
//   dx_j   += 1/d_i * sum_h q_ijh dZ_i[h,:]                 dP_jh += dl_ijh ;  dP_ih -= dl_ijh ;  dc_h += dl_ijh
// One warp per target node i (same traversal as the forward), lanes own channels; the per-edge 9 dot products are
// warp-reduced, scatter targets (dx_j, dP_j) are updated with red.global.add.f32.  The dense parts of the layer's backward
// (dZ = g.W_flat, dW = g^T.Z, dX += dP.U, dU = dP^T.X) are plain library GEMMs on the host side.
#include "common.cuh"

namespace geobi {

constexpr int H = GEOBI_HEADS;

template <int CPL>
__global__ void __launch_bounds__(256) feast_bwd_edges_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int C,
                                                              const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                              const double* __restrict__ P, const float* __restrict__ cvec,
                                                              const float* __restrict__ dZ, int64_t lddz, float* __restrict__ dx,
                                                              int64_t lddx, float* __restrict__ dP, float* __restrict__ dc) {
  __shared__ __align__(16) float qs[8][32][12];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int64_t i = (int64_t)blockIdx.x * 8 + warp;
  if (i >= N) return;
  const int b = rowptr[i];
  const int total = rowptr[i + 1] - b + 1;
  const float rcnt = 1.0f / (float)total;
  double Pi[H];
  float ch[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    Pi[h] = P[i * H + h];
    ch[h] = cvec[h];
  }
  // dZ_i slice owned by this lane: channels lane + 32k
  float dz[H][CPL];
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int k = 0; k < CPL; ++k) {
      const int c = lane + 32 * k;
      dz[h][k] = c < C ? dZ[i * lddz + h * C + c] * rcnt : 0.f;
    }
  float dPi[H], dcl[H];
#pragma unroll
  for (int h = 0; h < H; ++h) dPi[h] = dcl[h] = 0.f;

  for (int s0 = 0; s0 < total; s0 += 32) {
    const int s = s0 + lane;
    int j = (int)i;
    if (s < total) {
      if (s > 0) j = nbr[b + s - 1];
      float l[H];
      float m = -INFINITY;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = p_diff(P[(int64_t)j * H + h], Pi[h]) + ch[h];
        m = fmaxf(m, l[h]);
      }
      float sum = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        l[h] = expf(l[h] - m);
        sum += l[h];
      }
#pragma unroll
      for (int h = 0; h < H; ++h) qs[warp][lane][h] = l[h] / sum;
    }
    __syncwarp();
    const int cnt = min(32, total - s0);
    for (int t = 0; t < cnt; ++t) {
      const int jt = __shfl_sync(0xffffffffu, j, t);
      float q[H], dq[H];
#pragma unroll
      for (int h = 0; h < H; ++h) {
        q[h] = qs[warp][t][h];
        dq[h] = 0.f;
      }
      float gx[CPL];
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int c = lane + 32 * k;
        const float xv = c < C ? x[(int64_t)jt * ldx + c] : 0.f;
        float acc = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          dq[h] = fmaf(dz[h][k], xv, dq[h]);
          acc = fmaf(q[h], dz[h][k], acc);
        }
        gx[k] = acc;
      }
#pragma unroll
      for (int k = 0; k < CPL; ++k) {
        const int c = lane + 32 * k;
        if (c < C) atomicAdd(dx + (int64_t)jt * lddx + c, gx[k]);
      }
      float dot = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        dq[h] = warp_sum(dq[h]);
        dot = fmaf(q[h], dq[h], dot);
      }
      // lane h (< 9) finalises head h of this edge
      float dl = 0.f;
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float v = q[h] * (dq[h] - dot);
        if (lane == h) dl = v;
        dPi[h] -= v;
        dcl[h] += v;
      }
      if (lane < H) atomicAdd(dP + (int64_t)jt * H + lane, dl);
    }
    __syncwarp();
  }
  // every lane holds identical dPi / dcl: lane h writes head h
#pragma unroll
  for (int h = 0; h < H; ++h)
    if (lane == h) {
      atomicAdd(dP + i * H + h, dPi[h]);
      atomicAdd(dc + h, dcl[h]);
    }
}

// Round 2: the same backward for C in {32, 64, 128} (every layer but the two input layers) with
//   * LPR = C/4 lanes per row, each owning 4 adjacent channels: 32/LPR edges of a node in flight per warp instruction, the x_j row read
//     as one 128-bit load per lane, dx_j updated with ONE red.global.add.v4.f32 per lane (was 2-4 scalar atomics per lane and edge),
//     the 9 per-edge dot products reduced over LPR lanes (3-5 shuffles each, was 5 over the whole warp for one edge);
//   * persistent warps (grid = resident CTAs, nodes strided): dc is accumulated in registers across a warp's nodes and leaves the
//     kernel as 9 atomics per CTA - the one-warp-per-node form issued 9 atomics per NODE on the same 9 addresses.
__device__ __forceinline__ void red_add_v4(float* addr, float a, float b, float c, float d) {
  asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(addr), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
template <int LPR, int CH>
__global__ void __launch_bounds__(256) feast_bwd_edges_vec_kernel(const float* __restrict__ x, int64_t ldx, int64_t N, int C_rt,
                                                                  const int* __restrict__ rowptr, const int* __restrict__ nbr,
                                                                  const double* __restrict__ P, const float* __restrict__ cvec,
                                                                  const float* __restrict__ dZ, int64_t lddz, float* __restrict__ dx,
                                                                  int64_t lddx, float* __restrict__ dP, float* __restrict__ dc) {
  constexpr int EPW = 32 / LPR;
  const int C = CH == 4 ? 4 * LPR : C_rt;        // CH = 1: one channel per lane, lanes sl >= C idle (the 6 / 12-channel input layers)
  __shared__ __align__(16) float qs[8][32][12];
  __shared__ float dcs[8][H];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int eg = lane / LPR, sl = lane % LPR, c0 = CH * sl;
  const bool has_c = c0 < C;
  float ch[H], dcl[H];
#pragma unroll
  for (int h = 0; h < H; ++h) {
    ch[h] = cvec[h];
    dcl[h] = 0.f;
  }
  const int64_t n_warps = (int64_t)gridDim.x * 8;
  for (int64_t i = (int64_t)blockIdx.x * 8 + warp; i < N; i += n_warps) {
    const int b = rowptr[i];
    const int total = rowptr[i + 1] - b + 1;
    const float rcnt = 1.0f / (float)total;
    double Pi[H];
#pragma unroll
    for (int h = 0; h < H; ++h) Pi[h] = P[i * H + h];
    // dZ_i slice owned by this lane: channels c0 .. c0 + 3 of every head
    float4 dz[H];
#pragma unroll
    for (int h = 0; h < H; ++h) {
      if (CH == 4) {
        const float4 v = *reinterpret_cast<const float4*>(dZ + i * lddz + h * C + c0);
        dz[h] = make_float4(v.x * rcnt, v.y * rcnt, v.z * rcnt, v.w * rcnt);
      } else {
        dz[h] = make_float4(has_c ? dZ[i * lddz + h * C + c0] * rcnt : 0.f, 0.f, 0.f, 0.f);
      }
    }
    float dPi[H];
#pragma unroll
    for (int h = 0; h < H; ++h) dPi[h] = 0.f;
    for (int s0 = 0; s0 < total; s0 += 32) {
      const int s = s0 + lane;
      int j = (int)i;
      if (s < total) {
        if (s > 0) j = nbr[b + s - 1];
        float l[H];
        float m = -INFINITY;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          l[h] = p_diff(P[(int64_t)j * H + h], Pi[h]) + ch[h];
          m = fmaxf(m, l[h]);
        }
        float sum = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          l[h] = expf(l[h] - m);
          sum += l[h];
        }
#pragma unroll
        for (int h = 0; h < H; ++h) qs[warp][lane][h] = l[h] / sum;
      }
      __syncwarp();
      const int cnt = min(32, total - s0);
      for (int t0 = 0; t0 < cnt; t0 += EPW) {
        const int t = t0 + eg;
        const bool on = t < cnt;
        const int jt = __shfl_sync(0xffffffffu, j, on ? t : 0);
        float q[H];
#pragma unroll
        for (int h = 0; h < H; ++h) q[h] = on ? qs[warp][t][h] : 0.f;
        float4 xv = make_float4(0.f, 0.f, 0.f, 0.f);
        if (CH == 4) {
          if (on) xv = *reinterpret_cast<const float4*>(x + (int64_t)jt * ldx + c0);
        } else if (on && has_c) {
          xv.x = x[(int64_t)jt * ldx + c0];
        }
        float dq[H];
        float4 gx = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          if (CH == 4) {
            dq[h] = fmaf(dz[h].w, xv.w, fmaf(dz[h].z, xv.z, fmaf(dz[h].y, xv.y, dz[h].x * xv.x)));
            gx.y = fmaf(q[h], dz[h].y, gx.y);
            gx.z = fmaf(q[h], dz[h].z, gx.z);
            gx.w = fmaf(q[h], dz[h].w, gx.w);
          } else {
            dq[h] = dz[h].x * xv.x;
          }
          gx.x = fmaf(q[h], dz[h].x, gx.x);
        }
        if (dx != nullptr) {        // null: the layer's input needs no gradient (the network's two input layers)
          if (CH == 4) {
            if (on) red_add_v4(dx + (int64_t)jt * lddx + c0, gx.x, gx.y, gx.z, gx.w);
          } else if (on && has_c) {
            atomicAdd(dx + (int64_t)jt * lddx + c0, gx.x);
          }
        }
        float dot = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
#pragma unroll
          for (int o = LPR / 2; o > 0; o >>= 1) dq[h] += __shfl_xor_sync(0xffffffffu, dq[h], o);
          dot = fmaf(q[h], dq[h], dot);
        }
        // lane sl of the edge's group finalises head sl (and head sl + LPR where a group has fewer lanes than there are heads)
        float dl = 0.f, dl2 = 0.f;
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const float v = q[h] * (dq[h] - dot);      // 0 for an idle group (q = 0)
          if (sl == h) dl = v;
          if (LPR < H && sl + LPR == h) dl2 = v;
          dPi[h] -= v;
          dcl[h] += v;
        }
        if (on && sl < H) atomicAdd(dP + (int64_t)jt * H + sl, dl);
        if (LPR < H && on && sl + LPR < H) atomicAdd(dP + (int64_t)jt * H + sl + LPR, dl2);
      }
      __syncwarp();
    }
    // every lane of a group holds that group's share of dPi: add the groups, lane h (< 9) of group 0 writes head h
#pragma unroll
    for (int h = 0; h < H; ++h) {
      float v = dPi[h];
#pragma unroll
      for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == h) atomicAdd(dP + i * H + h, v);
    }
  }
  // dc: groups of a warp, then the warps of the CTA, then 9 atomics
#pragma unroll
  for (int h = 0; h < H; ++h) {
    float v = dcl[h];
#pragma unroll
    for (int o = LPR; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if (lane == 0) dcs[warp][h] = v;
  }
  __syncthreads();
  if (threadIdx.x < H) {
    float v = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) v += dcs[w][threadIdx.x];
    atomicAdd(dc + threadIdx.x, v);
  }
}

// dx[member argmax of segment s, c] = g[s, c]; every node belongs to exactly one segment, dx is zero-initialised.
__global__ void __launch_bounds__(256) segment_max_bwd_kernel(const float* __restrict__ x, int64_t ldx, int C, const int* __restrict__ rowptr,
                                                              const int* __restrict__ idx, int64_t n_seg, const float* __restrict__ g,
                                                              int64_t ldg, float* __restrict__ dx, int64_t lddx) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = t / C;
  if (s >= n_seg) return;
  const int c = (int)(t - s * C);
  const int b = rowptr[s], e = rowptr[s + 1];
  if (e <= b) return;
  int best = idx[b];
  float bv = x[(int64_t)best * ldx + c];
  for (int k = b + 1; k < e; ++k) {
    const int m = idx[k];
    const float v = x[(int64_t)m * ldx + c];
    if (v > bv) { bv = v; best = m; }   // ties: first (lowest id) member
  }
  dx[(int64_t)best * lddx + c] = g[s * ldg + c];
}

int feast_project_and_aggregate(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                const int32_t* row_map, int64_t n_src, const float* U, const float* c, double* P, void* Z, int64_t ldz,
                                int out_mode, cudaStream_t st);  // feast.cu
}  // namespace geobi

using namespace geobi;

// Z (fp32 [N, 9*C_in]) and P (fp64 [N, 9]) of the forward, for the backward pass.
extern "C" int geobi_feast_aggregate(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                     const float* U, const float* c, double* P, float* Z, void* stream) {
  GEOBI_REQUIRE(x && rowptr && U && c && P && Z && N >= 0 && c_in >= 1 && c_in <= 128, "feast_aggregate: bad arguments");
  if (N == 0) return GEOBI_OK;
  return feast_project_and_aggregate(x, ldx, N, c_in, rowptr, nbr, nullptr, N, U, c, P, Z, (int64_t)H * c_in, 0, static_cast<cudaStream_t>(stream));
}

namespace geobi {
// dZ has row stride lddz (>= 9*c_in; a multiple of 4 for the vector kernels): geobi_feast_bwd hands over its kpad-strided buffer
int feast_bwd_edges_launch(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr, const double* P,
                           const float* c, const float* dZ, int64_t lddz, float* dx, int64_t lddx, float* dP, float* dc, cudaStream_t st) {
  const unsigned blocks = (unsigned)cdiv(N, 8);
  const bool vec = (c_in == 32 || c_in == 64 || c_in == 128) && ldx % 4 == 0 && lddx % 4 == 0 && lddz % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(dx) & 15) == 0 && (reinterpret_cast<uintptr_t>(dZ) & 15) == 0;
  if (vec || c_in <= 16) {
    static int sms = 0;
    if (sms == 0) {
      int dev = 0;
      GEOBI_CUDA_OK(cudaGetDevice(&dev));
      GEOBI_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    }
    const unsigned grid = blocks < (unsigned)(sms * 6) ? blocks : (unsigned)(sms * 6);
#define BWD_ARGS x, ldx, N, c_in, rowptr, nbr, P, c, dZ, lddz, dx, lddx, dP, dc
    if (vec && c_in == 32) feast_bwd_edges_vec_kernel<8, 4><<<grid, 256, 0, st>>>(BWD_ARGS);
    else if (vec && c_in == 64) feast_bwd_edges_vec_kernel<16, 4><<<grid, 256, 0, st>>>(BWD_ARGS);
    else if (vec) feast_bwd_edges_vec_kernel<32, 4><<<grid, 256, 0, st>>>(BWD_ARGS);
    else if (c_in <= 8) feast_bwd_edges_vec_kernel<8, 1><<<grid, 256, 0, st>>>(BWD_ARGS);
    else feast_bwd_edges_vec_kernel<16, 1><<<grid, 256, 0, st>>>(BWD_ARGS);
#undef BWD_ARGS
    GEOBI_LAUNCH_OK("feast_bwd_edges (vec)");
    return GEOBI_OK;
  }
  GEOBI_REQUIRE(dx != nullptr, "feast_bwd_edges: dx may only be null for c_in <= 16, 32, 64 or 128");
  if (c_in <= 32) feast_bwd_edges_kernel<1><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, dZ, lddz, dx, lddx, dP, dc);
  else if (c_in <= 64) feast_bwd_edges_kernel<2><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, dZ, lddz, dx, lddx, dP, dc);
  else feast_bwd_edges_kernel<4><<<blocks, 256, 0, st>>>(x, ldx, N, c_in, rowptr, nbr, P, c, dZ, lddz, dx, lddx, dP, dc);
  GEOBI_LAUNCH_OK("feast_bwd_edges");
  return GEOBI_OK;
}
}  // namespace geobi

extern "C" int geobi_feast_bwd_edges(const float* x, int64_t ldx, int64_t N, int c_in, const int32_t* rowptr, const int32_t* nbr,
                                     const double* P, const float* c, const float* dZ, float* dx, int64_t lddx, float* dP, float* dc,
                                     void* stream) {
  GEOBI_REQUIRE(x && rowptr && P && c && dZ && dP && dc && N >= 0 && c_in >= 1 && c_in <= 128, "feast_bwd_edges: bad arguments");
  if (N == 0) return GEOBI_OK;
  return feast_bwd_edges_launch(x, ldx, N, c_in, rowptr, nbr, P, c, dZ, (int64_t)H * c_in, dx, lddx, dP, dc, static_cast<cudaStream_t>(stream));
}

extern "C" int geobi_segment_max_bwd(const float* x, int64_t ldx, int channels, const int32_t* rowptr, const int32_t* idx, int64_t n_seg,
                                     const float* g, int64_t ldg, float* dx, int64_t lddx, void* stream) {
  GEOBI_REQUIRE(x && rowptr && idx && g && dx && channels > 0 && n_seg >= 0, "segment_max_bwd: bad arguments");
  if (n_seg == 0) return GEOBI_OK;
  segment_max_bwd_kernel<<<(unsigned)cdiv(n_seg * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, channels, rowptr, idx,
                                                                                                               n_seg, g, ldg, dx, lddx);
  GEOBI_LAUNCH_OK("segment_max_bwd");
  return GEOBI_OK;
}
