// Segment-reduce / gather / geometric kernels of libgeobi: cluster pooling, unpooling,
// dual-domain transfer (vertex<->facet), bilateral and feature edge weights, vertex update.
// All are HBM/L2-bound streaming kernels: coalesced row access, one pass, no atomics.
#include "common.cuh"

namespace geobi {

// ------------------------------------------------------------------------------ segment reduce
// thread = (segment, channel): consecutive threads read consecutive floats of a gathered row.
__global__ void __launch_bounds__(256) segment_reduce_kernel(const float* __restrict__ x, int64_t ldx, int C, const int* __restrict__ rowptr,
                                                             const int* __restrict__ idx, int fixed, int64_t n_seg, int op,
                                                             float* __restrict__ out, int64_t ldo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = t / C;
  if (s >= n_seg) return;
  const int c = (int)(t - s * C);
  const int64_t b = rowptr ? (int64_t)rowptr[s] : s * (int64_t)fixed;
  const int64_t e = rowptr ? (int64_t)rowptr[s + 1] : b + fixed;
  float acc = 0.f;
  if (op == 1) {
    if (e > b) {
      acc = x[(int64_t)idx[b] * ldx + c];
      for (int64_t k = b + 1; k < e; ++k) acc = fmaxf(acc, x[(int64_t)idx[k] * ldx + c]);
    }
  } else {
    for (int64_t k = b; k < e; ++k) acc += x[(int64_t)idx[k] * ldx + c];
    if (op == 0) {
      const float cnt = (float)(e - b > 1 ? e - b : 1);
      acc = acc / cnt;
    }
  }
  out[s * ldo + c] = acc;
}

// float4 variant (C, ldx, ldo multiples of 4, 16-byte aligned bases): thread = (segment, 4 channels)
__global__ void __launch_bounds__(256) segment_reduce_vec_kernel(const float* __restrict__ x, int64_t ldx, int C4, const int* __restrict__ rowptr,
                                                                 const int* __restrict__ idx, int fixed, int64_t n_seg, int op,
                                                                 float* __restrict__ out, int64_t ldo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t s = t / C4;
  if (s >= n_seg) return;
  const int c = (int)(t - s * C4) * 4;
  const int64_t b = rowptr ? (int64_t)rowptr[s] : s * (int64_t)fixed;
  const int64_t e = rowptr ? (int64_t)rowptr[s + 1] : b + fixed;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  if (op == 1) {
    if (e > b) {
      acc = *reinterpret_cast<const float4*>(x + (int64_t)idx[b] * ldx + c);
      for (int64_t k = b + 1; k < e; ++k) {
        const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)idx[k] * ldx + c);
        acc.x = fmaxf(acc.x, v.x); acc.y = fmaxf(acc.y, v.y); acc.z = fmaxf(acc.z, v.z); acc.w = fmaxf(acc.w, v.w);
      }
    }
  } else {
    for (int64_t k = b; k < e; ++k) {
      const float4 v = *reinterpret_cast<const float4*>(x + (int64_t)idx[k] * ldx + c);
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
    if (op == 0) {
      const float cnt = (float)(e - b > 1 ? e - b : 1);
      acc.x = acc.x / cnt; acc.y = acc.y / cnt; acc.z = acc.z / cnt; acc.w = acc.w / cnt;
    }
  }
  *reinterpret_cast<float4*>(out + s * ldo + c) = acc;
}

__global__ void __launch_bounds__(256) gather_rows_kernel(const float* __restrict__ x, int64_t ldx, int C, const int* __restrict__ idx,
                                                          int64_t n_out, float* __restrict__ out, int64_t ldo) {
  const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t i = t / C;
  if (i >= n_out) return;
  const int c = (int)(t - i * C);
  out[i * ldo + c] = x[(int64_t)idx[i] * ldx + c];
}

// ------------------------------------------------------------------------------ feature edge weight
// warp per row i; per CSR entry: squared feature distance by a lane-strided dot + shuffle reduce.
__global__ void __launch_bounds__(256) edge_weight_feat_kernel(const float* __restrict__ x, int64_t ldx, int C, const int* __restrict__ rowptr,
                                                               const int* __restrict__ nbr, int64_t n, const float* __restrict__ w_in,
                                                               int mode, float param, float* __restrict__ w_out) {
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int b = rowptr[i], e = rowptr[i + 1];
  float xi[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const int c = lane + 32 * k;
    xi[k] = c < C ? x[i * ldx + c] : 0.f;
  }
  for (int q = b; q < e; ++q) {
    const int64_t j = nbr[q];
    float d2 = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int c = lane + 32 * k;
      if (c < C) {
        const float d = xi[k] - x[j * ldx + c];
        d2 += d * d;
      }
    }
    d2 = warp_sum(d2);
    if (lane == 0) {
      float r;
      if (mode == 0) r = d2;
      else if (mode == 1) r = expf(d2 / (-param));
      else if (mode == 2) r = w_in[q] * expf(d2 / (-param));
      else r = w_in[q] + expf(d2 / (-2.0f));
      w_out[q] = r;
    }
  }
}

// Vectorised variant for C in {32, 64, 128}: LPE = C/4 lanes per edge (one 128-bit load each), 32/LPE edges per warp
// instruction, log2(LPE) shuffles per edge group instead of 5 per edge.
template <int LPE>
__global__ void __launch_bounds__(256) edge_weight_feat_vec_kernel(const float* __restrict__ x, int64_t ldx, const int* __restrict__ rowptr,
                                                                   const int* __restrict__ nbr, int64_t n, const float* __restrict__ w_in,
                                                                   int mode, float param, float* __restrict__ w_out) {
  constexpr int EPW = 32 / LPE;   // edges per warp instruction
  const int lane = threadIdx.x & 31;
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (i >= n) return;
  const int eg = lane / LPE, sl = lane % LPE;
  const int b = rowptr[i], e = rowptr[i + 1];
  const float4 xi = *reinterpret_cast<const float4*>(x + i * ldx + sl * 4);
  for (int q0 = b; q0 < e; q0 += EPW) {
    const int q = q0 + eg;
    float d2 = 0.f;
    if (q < e) {
      const int64_t j = nbr[q];
      const float4 xj = *reinterpret_cast<const float4*>(x + j * ldx + sl * 4);
      const float dx = xi.x - xj.x, dy = xi.y - xj.y, dz = xi.z - xj.z, dw = xi.w - xj.w;
      d2 = dx * dx + dy * dy + dz * dz + dw * dw;
    }
#pragma unroll
    for (int o = LPE / 2; o > 0; o >>= 1) d2 += __shfl_xor_sync(0xffffffffu, d2, o);
    if (sl == 0 && q < e) {
      float r;
      if (mode == 0) r = d2;
      else if (mode == 1) r = expf(d2 / (-param));
      else if (mode == 2) r = w_in[q] * expf(d2 / (-param));
      else r = w_in[q] + expf(d2 / (-2.0f));
      w_out[q] = r;
    }
  }
}

// ------------------------------------------------------------------------------ calc_weight
__device__ __forceinline__ float edge_l2(const float* __restrict__ pos, int64_t a, int64_t b) {
  const float dx = pos[a * 3] - pos[b * 3], dy = pos[a * 3 + 1] - pos[b * 3 + 1], dz = pos[a * 3 + 2] - pos[b * 3 + 2];
  return dx * dx + dy * dy + dz * dz;
}
__global__ void __launch_bounds__(256) edge_len_partial_kernel(const float* __restrict__ pos, const int64_t* __restrict__ row,
                                                               const int64_t* __restrict__ col, int64_t E, double* __restrict__ partial) {
  double s = 0.0;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x)
    s += (double)sqrtf(edge_l2(pos, row[e], col[e]));
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) edge_len_final_kernel(double* partial, int nb, int64_t E) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[nb] = sh[0] / (double)(E > 0 ? E : 1);
}
__global__ void __launch_bounds__(256) calc_weight_kernel(const float* __restrict__ pos, const float* __restrict__ nrm,
                                                          const int64_t* __restrict__ row, const int64_t* __restrict__ col, int64_t E,
                                                          const double* __restrict__ mean_len, float* __restrict__ w_out) {
  const float denom = -2.0f * (float)(*mean_len) + 1e-12f;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < E; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t a = row[e], b = col[e];
    const float l2 = edge_l2(pos, a, b);
    float dn = nrm[a * 3] * nrm[b * 3];
    dn += nrm[a * 3 + 1] * nrm[b * 3 + 1];
    dn += nrm[a * 3 + 2] * nrm[b * 3 + 2];
    w_out[e] = fmaxf(dn, 0.001f) * expf(l2 / denom);
  }
}

// The same weights for a loop-free CSR (rows = sources), in CSR entry order: what the matcher reads.  The mean edge length is
// taken over the reference's list, i.e. the CSR entries PLUS `n_loops` zero-length self loops (to_undirected_with_self_loops /
// the facet builder's self entries).  8 lanes per row.
__global__ void __launch_bounds__(256) edge_len_partial_csr_kernel(const float* __restrict__ pos, const int* __restrict__ rowptr,
                                                                   const int* __restrict__ nbr, int64_t n, double* __restrict__ partial) {
  double s = 0.0;
  const int sub = threadIdx.x & 7;
  for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; i < n; i += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int b = rowptr[i], e = rowptr[i + 1];
    for (int q = b + sub; q < e; q += 8) s += (double)sqrtf(edge_l2(pos, i, nbr[q]));
  }
  __shared__ double sh[256];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) partial[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) edge_len_final_csr_kernel(double* partial, int nb, const int* __restrict__ rowptr, int64_t n, int64_t n_loops) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < nb; i += 256) s += partial[i];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const int64_t E = (int64_t)rowptr[n] + n_loops;
    partial[nb] = sh[0] / (double)(E > 0 ? E : 1);
  }
}
__global__ void __launch_bounds__(256) calc_weight_csr_kernel(const float* __restrict__ pos, const float* __restrict__ nrm, const int* __restrict__ rowptr,
                                                              const int* __restrict__ nbr, int64_t n, const double* __restrict__ mean_len,
                                                              float* __restrict__ w_out) {
  const float denom = -2.0f * (float)(*mean_len) + 1e-12f;
  const int sub = threadIdx.x & 7;
  for (int64_t a = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; a < n; a += ((int64_t)gridDim.x * blockDim.x) >> 3) {
    const int b0 = rowptr[a], e0 = rowptr[a + 1];
    for (int q = b0 + sub; q < e0; q += 8) {
      const int64_t b = nbr[q];
      const float l2 = edge_l2(pos, a, b);
      float dn = nrm[a * 3] * nrm[b * 3];
      dn += nrm[a * 3 + 1] * nrm[b * 3 + 1];
      dn += nrm[a * 3 + 2] * nrm[b * 3 + 2];
      w_out[q] = fmaxf(dn, 0.001f) * expf(l2 / denom);
    }
  }
}

// ------------------------------------------------------------------------------ face normal / v->f transfer
__device__ __forceinline__ void face_geom(const float* __restrict__ p, int64_t ldp, const int64_t* __restrict__ fv, int64_t f, float* cent,
                                          float* nrm) {
  const int64_t a = fv[f * 3], b = fv[f * 3 + 1], c = fv[f * 3 + 2];
  const float ax = p[a * ldp], ay = p[a * ldp + 1], az = p[a * ldp + 2];
  const float bx = p[b * ldp], by = p[b * ldp + 1], bz = p[b * ldp + 2];
  const float cx = p[c * ldp], cy = p[c * ldp + 1], cz = p[c * ldp + 2];
  if (cent) {
    cent[0] = (ax + bx + cx) / 3.0f;
    cent[1] = (ay + by + cy) / 3.0f;
    cent[2] = (az + bz + cz) / 3.0f;
  }
  const float ux = bx - ax, uy = by - ay, uz = bz - az;
  const float vx = cx - ax, vy = cy - ay, vz = cz - az;
  // un-fused products, as the eager reference evaluates torch.cross
  const float nx = __fsub_rn(__fmul_rn(uy, vz), __fmul_rn(uz, vy));
  const float ny = __fsub_rn(__fmul_rn(uz, vx), __fmul_rn(ux, vz));
  const float nz = __fsub_rn(__fmul_rn(ux, vy), __fmul_rn(uy, vx));
  const float len = sqrtf(__fadd_rn(__fadd_rn(__fmul_rn(nx, nx), __fmul_rn(ny, ny)), __fmul_rn(nz, nz)));
  const float d = fmaxf(len, 1e-12f);
  nrm[0] = nx / d;
  nrm[1] = ny / d;
  nrm[2] = nz / d;
}
__global__ void __launch_bounds__(256) face_normal_kernel(const float* __restrict__ p, int64_t ldp, const int64_t* __restrict__ fv, int64_t F,
                                                          float* __restrict__ out, int64_t ldo) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float n[3];
  face_geom(p, ldp, fv, f, nullptr, n);
  out[f * ldo] = n[0];
  out[f * ldo + 1] = n[1];
  out[f * ldo + 2] = n[2];
}
__global__ void __launch_bounds__(256) v2f_transfer_kernel(const float* __restrict__ p, int64_t ldp, const int64_t* __restrict__ fv,
                                                           const float* __restrict__ xf, int64_t ldxf, int cf, int64_t F,
                                                           float* __restrict__ out, int64_t ldo) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  float c[3], n[3];
  face_geom(p, ldp, fv, f, c, n);
  float* o = out + f * ldo;
  for (int k = 0; k < cf; ++k) o[k] = xf[f * ldxf + k];
  o[cf] = c[0]; o[cf + 1] = c[1]; o[cf + 2] = c[2];
  o[cf + 3] = n[0]; o[cf + 4] = n[1]; o[cf + 5] = n[2];
}

// Backward of the transfer (training step): g = d loss / d out[:, cf:cf+6] = (gradient of the corner mean | gradient of the unit normal).
//   centroid: each corner gets g_c / 3;  normal n = u / max(|u|, eps), u = (p1 - p0) x (p2 - p0):
//   g_u = (g_n - n <n, g_n>) / |u|  (g_n / eps below the clamp),  g_e1 = e2 x g_u,  g_e2 = g_u x e1,  p1 += g_e1, p2 += g_e2, p0 -= g_e1 + g_e2.
// Vertices collect their faces' contributions with atomics (6 faces per vertex on average); d_feat_v is zero-initialised by the caller.
__global__ void __launch_bounds__(256) v2f_transfer_bwd_kernel(const float* __restrict__ p, int64_t ldp, const int64_t* __restrict__ fv,
                                                               const float* __restrict__ g, int64_t ldg, int64_t F, float* __restrict__ dp,
                                                               int64_t lddp, bool normals_only) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int64_t a = fv[f * 3], b = fv[f * 3 + 1], c = fv[f * 3 + 2];
  const float ax = p[a * ldp], ay = p[a * ldp + 1], az = p[a * ldp + 2];
  const float e1x = p[b * ldp] - ax, e1y = p[b * ldp + 1] - ay, e1z = p[b * ldp + 2] - az;
  const float e2x = p[c * ldp] - ax, e2y = p[c * ldp + 1] - ay, e2z = p[c * ldp + 2] - az;
  const float ux = e1y * e2z - e1z * e2y, uy = e1z * e2x - e1x * e2z, uz = e1x * e2y - e1y * e2x;
  const float len = sqrtf(ux * ux + uy * uy + uz * uz);
  const float* gr = g + f * ldg;
  // normals_only (geobi_face_normal_bwd): a row of g is the normal's gradient alone
  const float gcx = normals_only ? 0.f : gr[0] / 3.0f, gcy = normals_only ? 0.f : gr[1] / 3.0f, gcz = normals_only ? 0.f : gr[2] / 3.0f;
  const float gnx = normals_only ? gr[0] : gr[3], gny = normals_only ? gr[1] : gr[4], gnz = normals_only ? gr[2] : gr[5];
  float gux, guy, guz;
  if (len > 1e-12f) {
    const float nx = ux / len, ny = uy / len, nz = uz / len;
    const float dot = nx * gnx + ny * gny + nz * gnz;
    gux = (gnx - nx * dot) / len;
    guy = (gny - ny * dot) / len;
    guz = (gnz - nz * dot) / len;
  } else {                                  // below normalize's clamp the output is u / eps
    gux = gnx / 1e-12f;
    guy = gny / 1e-12f;
    guz = gnz / 1e-12f;
  }
  // u = e1 x e2
  const float g1x = e2y * guz - e2z * guy, g1y = e2z * gux - e2x * guz, g1z = e2x * guy - e2y * gux;   // e2 x g_u
  const float g2x = guy * e1z - guz * e1y, g2y = guz * e1x - gux * e1z, g2z = gux * e1y - guy * e1x;   // g_u x e1
  atomicAdd(dp + b * lddp, gcx + g1x);
  atomicAdd(dp + b * lddp + 1, gcy + g1y);
  atomicAdd(dp + b * lddp + 2, gcz + g1z);
  atomicAdd(dp + c * lddp, gcx + g2x);
  atomicAdd(dp + c * lddp + 1, gcy + g2y);
  atomicAdd(dp + c * lddp + 2, gcz + g2z);
  atomicAdd(dp + a * lddp, gcx - g1x - g2x);
  atomicAdd(dp + a * lddp + 1, gcy - g1y - g2y);
  atomicAdd(dp + a * lddp + 2, gcz - g1z - g2z);
}

// ------------------------------------------------------------------------------ vertex update
__global__ void __launch_bounds__(256) face_centroid_kernel(const float* __restrict__ p, const int64_t* __restrict__ fv, int64_t F,
                                                            float* __restrict__ cent) {
  const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= F) return;
  const int64_t a = fv[f * 3], b = fv[f * 3 + 1], c = fv[f * 3 + 2];
#pragma unroll
  for (int k = 0; k < 3; ++k) cent[f * 3 + k] = (p[a * 3 + k] + p[b * 3 + k] + p[c * 3 + k]) / 3.0f;
}
__global__ void __launch_bounds__(256) vertex_update_kernel(const float* __restrict__ p, const float* __restrict__ cent,
                                                            const int64_t* __restrict__ vf, int K, const float* __restrict__ fn,
                                                            const float* __restrict__ depth, int64_t V, float* __restrict__ q) {
  const int64_t v = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= V) return;
  const float px = p[v * 3], py = p[v * 3 + 1], pz = p[v * 3 + 2];
  float ax = 0.f, ay = 0.f, az = 0.f;
  int cnt = 0;
  for (int k = 0; k < K; ++k) {
    const int64_t f = vf[v * K + k];
    if (f < 0) continue;
    ++cnt;
    const float nx = fn[f * 3], ny = fn[f * 3 + 1], nz = fn[f * 3 + 2];
    const float d = nx * (cent[f * 3] - px) + ny * (cent[f * 3 + 1] - py) + nz * (cent[f * 3 + 2] - pz);
    ax += nx * d; ay += ny * d; az += nz * d;
  }
  const float inv = (float)(cnt > 1 ? cnt : 1);
  ax /= inv; ay /= inv; az /= inv;
  if (depth) {
    const float dx = depth[v * 3], dy = depth[v * 3 + 1], dz = depth[v * 3 + 2];
    const float s = ax * dx + ay * dy + az * dz;
    ax = s * dx; ay = s * dy; az = s * dz;
  }
  q[v * 3] = px + ax; q[v * 3 + 1] = py + ay; q[v * 3 + 2] = pz + az;
}

}  // namespace geobi

using namespace geobi;

extern "C" int geobi_segment_reduce(const float* x, int64_t ldx, int channels, const int32_t* rowptr, const int32_t* idx, int fixed,
                                    int64_t n_seg, int op, float* out, int64_t ldo, void* stream) {
  GEOBI_REQUIRE(x && idx && out && channels > 0 && n_seg >= 0 && op >= 0 && op <= 2 && (rowptr || fixed > 0), "segment_reduce: bad arguments");
  if (n_seg == 0) return GEOBI_OK;
  const bool vec = channels % 4 == 0 && ldx % 4 == 0 && ldo % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 &&
                   (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (vec)
    segment_reduce_vec_kernel<<<(unsigned)cdiv(n_seg * (channels / 4), 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        x, ldx, channels / 4, rowptr, idx, fixed, n_seg, op, out, ldo);
  else
    segment_reduce_kernel<<<(unsigned)cdiv(n_seg * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, channels, rowptr, idx, fixed,
                                                                                                                n_seg, op, out, ldo);
  GEOBI_LAUNCH_OK("segment_reduce");
  return GEOBI_OK;
}

extern "C" int geobi_gather_rows(const float* x, int64_t ldx, int channels, const int32_t* idx, int64_t n_out, float* out, int64_t ldo,
                                 void* stream) {
  GEOBI_REQUIRE(x && idx && out && channels > 0 && n_out >= 0, "gather_rows: bad arguments");
  if (n_out == 0) return GEOBI_OK;
  gather_rows_kernel<<<(unsigned)cdiv(n_out * channels, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(x, ldx, channels, idx, n_out, out, ldo);
  GEOBI_LAUNCH_OK("gather_rows");
  return GEOBI_OK;
}

extern "C" int geobi_edge_weight_feat(const float* x, int64_t ldx, int channels, const int32_t* rowptr, const int32_t* nbr, int64_t n_nodes,
                                      const float* w_in, int mode, float param, float* w_out, void* stream) {
  GEOBI_REQUIRE(x && rowptr && w_out && n_nodes >= 0, "edge_weight_feat: bad arguments");
  GEOBI_REQUIRE(channels > 0 && channels <= 128, "edge_weight_feat: channels must be in 1..128 (got %d)", channels);
  GEOBI_REQUIRE(mode == 0 || mode == 1 || mode == 2 || mode == 10, "edge_weight_feat: unsupported mode %d", mode);
  GEOBI_REQUIRE(!(mode == 2 || mode == 10) || w_in, "edge_weight_feat: mode %d needs w_in", mode);
  if (n_nodes == 0) return GEOBI_OK;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned blocks = (unsigned)cdiv(n_nodes * 32, 256);
  const bool vec = (channels == 32 || channels == 64 || channels == 128) && ldx % 4 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0;
  if (vec && channels == 32) edge_weight_feat_vec_kernel<8><<<blocks, 256, 0, st>>>(x, ldx, rowptr, nbr, n_nodes, w_in, mode, param, w_out);
  else if (vec && channels == 64) edge_weight_feat_vec_kernel<16><<<blocks, 256, 0, st>>>(x, ldx, rowptr, nbr, n_nodes, w_in, mode, param, w_out);
  else if (vec) edge_weight_feat_vec_kernel<32><<<blocks, 256, 0, st>>>(x, ldx, rowptr, nbr, n_nodes, w_in, mode, param, w_out);
  else edge_weight_feat_kernel<<<blocks, 256, 0, st>>>(x, ldx, channels, rowptr, nbr, n_nodes, w_in, mode, param, w_out);
  GEOBI_LAUNCH_OK("edge_weight_feat");
  return GEOBI_OK;
}

static const int CW_BLOCKS = 1024;
extern "C" size_t geobi_calc_weight_ws_bytes(int64_t) { return align256((CW_BLOCKS + 2) * sizeof(double)) + 256; }

__global__ void mean_to_float_kernel(const double* __restrict__ mean, float* __restrict__ out) { *out = (float)(*mean); }
extern "C" int geobi_mean_edge_length_csr(const float* pos, const int32_t* rowptr, const int32_t* nbr, int64_t n_nodes, float* mean_out, void* ws,
                                          size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(pos && rowptr && nbr && mean_out && n_nodes > 0, "mean_edge_length_csr: bad arguments");
  if (!ws || ws_bytes < geobi_calc_weight_ws_bytes(n_nodes)) { set_error("mean_edge_length_csr: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  int nb = (int)cdiv(n_nodes * 8, 256);
  if (nb > CW_BLOCKS) nb = CW_BLOCKS;
  edge_len_partial_csr_kernel<<<nb, 256, 0, st>>>(pos, rowptr, nbr, n_nodes, partial);
  edge_len_final_csr_kernel<<<1, 256, 0, st>>>(partial, nb, rowptr, n_nodes, 0);
  mean_to_float_kernel<<<1, 1, 0, st>>>(partial + nb, mean_out);
  GEOBI_LAUNCH_OK("mean_edge_length_csr");
  return GEOBI_OK;
}

extern "C" int geobi_calc_weight_csr(const float* pos, const float* nrm, const int32_t* rowptr, const int32_t* nbr, int64_t n_nodes, int64_t n_loops,
                                     float* w_out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(pos && nrm && rowptr && nbr && w_out && n_nodes >= 0 && n_loops >= 0, "calc_weight_csr: bad arguments");
  if (n_nodes == 0) return GEOBI_OK;
  if (!ws || ws_bytes < geobi_calc_weight_ws_bytes(n_nodes)) { set_error("calc_weight_csr: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  int nb = (int)cdiv(n_nodes * 8, 256);
  if (nb > CW_BLOCKS) nb = CW_BLOCKS;
  edge_len_partial_csr_kernel<<<nb, 256, 0, st>>>(pos, rowptr, nbr, n_nodes, partial);
  edge_len_final_csr_kernel<<<1, 256, 0, st>>>(partial, nb, rowptr, n_nodes, n_loops);
  calc_weight_csr_kernel<<<nb, 256, 0, st>>>(pos, nrm, rowptr, nbr, n_nodes, partial + nb, w_out);
  GEOBI_LAUNCH_OK("calc_weight_csr");
  return GEOBI_OK;
}

extern "C" int geobi_calc_weight(const float* pos, const float* nrm, const int64_t* row, const int64_t* col, int64_t n_edges, float* w_out,
                                 void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(pos && nrm && row && col && w_out && n_edges >= 0, "calc_weight: bad arguments");
  if (n_edges == 0) return GEOBI_OK;
  if (!ws || ws_bytes < geobi_calc_weight_ws_bytes(n_edges)) { set_error("calc_weight: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  double* partial = static_cast<double*>(ws);
  int nb = (int)cdiv(n_edges, 256);
  if (nb > CW_BLOCKS) nb = CW_BLOCKS;
  edge_len_partial_kernel<<<nb, 256, 0, st>>>(pos, row, col, n_edges, partial);
  edge_len_final_kernel<<<1, 256, 0, st>>>(partial, nb, n_edges);
  calc_weight_kernel<<<nb, 256, 0, st>>>(pos, nrm, row, col, n_edges, partial + nb, w_out);
  GEOBI_LAUNCH_OK("calc_weight");
  return GEOBI_OK;
}

extern "C" int geobi_face_normal(const float* points, int64_t ldp, const int64_t* fv, int64_t n_faces, float* out, int64_t ldo, void* stream) {
  GEOBI_REQUIRE(points && fv && out && n_faces >= 0 && ldp >= 3 && ldo >= 3, "face_normal: bad arguments");
  if (n_faces == 0) return GEOBI_OK;
  face_normal_kernel<<<(unsigned)cdiv(n_faces, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(points, ldp, fv, n_faces, out, ldo);
  GEOBI_LAUNCH_OK("face_normal");
  return GEOBI_OK;
}

extern "C" int geobi_v2f_transfer(const float* feat_v, int64_t ldv, const int64_t* fv, const float* xf, int64_t ldxf, int cf, int64_t n_faces,
                                  float* out, int64_t ldo, void* stream) {
  GEOBI_REQUIRE(feat_v && fv && out && (xf || cf == 0) && cf >= 0 && n_faces >= 0 && ldv >= 3 && ldo >= cf + 6, "v2f_transfer: bad arguments");
  if (n_faces == 0) return GEOBI_OK;
  v2f_transfer_kernel<<<(unsigned)cdiv(n_faces, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat_v, ldv, fv, xf, ldxf, cf, n_faces, out, ldo);
  GEOBI_LAUNCH_OK("v2f_transfer");
  return GEOBI_OK;
}

extern "C" int geobi_v2f_transfer_bwd(const float* feat_v, int64_t ldv, const int64_t* fv, const float* g_out, int64_t ldg, int64_t n_faces,
                                      float* d_feat_v, int64_t lddv, void* stream) {
  GEOBI_REQUIRE(feat_v && fv && g_out && d_feat_v && n_faces >= 0 && ldv >= 3 && ldg >= 6 && lddv >= 3, "v2f_transfer_bwd: bad arguments");
  if (n_faces == 0) return GEOBI_OK;
  v2f_transfer_bwd_kernel<<<(unsigned)cdiv(n_faces, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(feat_v, ldv, fv, g_out, ldg, n_faces,
                                                                                                        d_feat_v, lddv, false);
  GEOBI_LAUNCH_OK("v2f_transfer_bwd");
  return GEOBI_OK;
}

extern "C" int geobi_face_normal_bwd(const float* points, int64_t ldp, const int64_t* fv, const float* g_normal, int64_t ldg, int64_t n_faces,
                                     float* d_points, int64_t lddp, void* stream) {
  GEOBI_REQUIRE(points && fv && g_normal && d_points && n_faces >= 0 && ldp >= 3 && ldg >= 3 && lddp >= 3, "face_normal_bwd: bad arguments");
  if (n_faces == 0) return GEOBI_OK;
  v2f_transfer_bwd_kernel<<<(unsigned)cdiv(n_faces, 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(points, ldp, fv, g_normal, ldg, n_faces,
                                                                                                        d_points, lddp, true);
  GEOBI_LAUNCH_OK("face_normal_bwd");
  return GEOBI_OK;
}

extern "C" size_t geobi_update_position_ws_bytes(int64_t n_verts, int64_t n_faces) {
  return align256((size_t)n_faces * 3 * sizeof(float)) + align256((size_t)n_verts * 3 * sizeof(float)) + 512;
}

extern "C" int geobi_update_position(const float* points, const int64_t* fv, const int64_t* vf, int64_t k, const float* face_normals, int n_iter,
                                     const float* depth, int64_t n_verts, int64_t n_faces, float* out, void* ws, size_t ws_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GEOBI_REQUIRE(points && fv && vf && face_normals && out && out != points && k > 0 && n_iter >= 0 && n_verts >= 0 && n_faces >= 0,
                "update_position: bad arguments");
  if (!ws || ws_bytes < geobi_update_position_ws_bytes(n_verts, n_faces)) { set_error("update_position: workspace too small"); return GEOBI_ERR_WORKSPACE; }
  Carver c(ws, ws_bytes);
  float* cent = c.take<float>((size_t)n_faces * 3);
  float* tmp = c.take<float>((size_t)n_verts * 3);
  if (n_verts == 0) return GEOBI_OK;
  if (n_iter == 0) {
    GEOBI_CUDA_OK(cudaMemcpyAsync(out, points, sizeof(float) * 3 * n_verts, cudaMemcpyDeviceToDevice, st));
    return GEOBI_OK;
  }
  // ping-pong so that the last sweep lands in `out`
  const float* src = points;
  for (int it = 0; it < n_iter; ++it) {
    float* dst = ((n_iter - 1 - it) % 2 == 0) ? out : tmp;
    if (n_faces > 0) face_centroid_kernel<<<(unsigned)cdiv(n_faces, 256), 256, 0, st>>>(src, fv, n_faces, cent);
    vertex_update_kernel<<<(unsigned)cdiv(n_verts, 256), 256, 0, st>>>(src, cent, vf, (int)k, face_normals, depth, n_verts, dst);
    src = dst;
  }
  GEOBI_LAUNCH_OK("update_position");
  return GEOBI_OK;
}
