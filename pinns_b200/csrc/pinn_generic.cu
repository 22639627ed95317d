// Generic FP32 kernel: any layer widths, Burgers (4 streams) / Euler (3 streams) /
// data term (1 stream).  One CTA owns a tile of 32 collocation points at a time and
// walks the network layer by layer; per-layer activations live in a CTA-private
// scratch slab (L2 resident), GEMM operands are register tiles:
//   F/B:  thread = (point, 8 outputs)  -> acc[S][8], weights via broadcast LDG.128
//   G  :  thread = (8x8 tile of W-bar) -> k-rows staged transposed in shared memory
// This is the correctness path for every configuration and the production path for
// nets the fused thread-per-point kernel (pinn_fused.cu) does not cover.
//
// Math: SURVEY.md appendix A.2/A.3, i.e. the Taylor-forward + single-reverse schedule of
// neural_net/net_u/net_f (INF-L2:96-120, AB-ADMM:170-180, EUL:176-198) and the loss
// variants INF-L2:68-69, INF-ADMM:98-100, ID-L2b:57-58, AB-L2:59-60, AB-ADMM:129-130, EUL:128-133.
#include <cstring>

#include "pinn_kernels.h"

namespace {

constexpr int T = PINN_TILE;
// Threads per CTA: 8 warps (two CTAs per SM at large N), or 9 where the ninth warp saves a whole round of the F / B GEMMs: a
// warp owns one 8-neuron output group per round, and the reference's width 200 is 25 groups -- 8 warps need 4 rounds (8, 8, 8, 1),
// a cluster of 3 CTAs two (24, 1); with 9 warps it is 3 rounds, or one.  Measured: Euler [2,200x5,3] at 1000 points 289 -> 257 us
// per Adam step, [2,200x8,1] 610 -> 527 us; for nets whose round count does not change 9 warps only cost the second CTA per SM
// ([2,128x8,1] at 65 536 points 6.5 -> 8.3 ms), so the choice is made per net (gen_threads_for).
constexpr int GEN_THREADS = 256;
constexpr int GEN_THREADS_WIDE = 288;

// A 32-point tile is owned by a cluster of `cs` CTAs (cs = 1 at large N; up to 8 when the job has fewer tiles than the
// GPU has CTA slots, e.g. the reference's N_f = 1000 batches of a 200-wide net = 32 tiles): the CTAs split the output
// neurons of every layer / the tiles of every weight gradient and meet at cluster barriers; activations travel through
// the tile's global scratch slab (barrier.cluster release/acquire orders the writes and invalidates L1).
__device__ __forceinline__ void tile_sync(int cs) {
  if (cs == 1) {
    __syncthreads();
  } else {
    asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
  }
}

#ifdef PINN_TRACE
// debug builds (-DPINN_TRACE): CTA 0 thread 0 logs (tag, clock64) at phase boundaries; read back with pinn_debug_trace
__device__ long long g_trace[2048];
__device__ int g_trace_n;
#define TRACE(tag)                                                              \
  do {                                                                          \
    if (blockIdx.x == 0 && threadIdx.x == 0 && g_trace_n < 1023) {                     \
      g_trace[2 * g_trace_n] = (tag);                                           \
      g_trace[2 * g_trace_n + 1] = clock64();                                   \
      ++g_trace_n;                                                              \
    }                                                                           \
  } while (0)
#else
#define TRACE(tag)
#endif

template <int S>
struct Streams {
  const float* p[S];
};

// input streams of weight layer l: the Taylor seeds for l == 0, else the previous hidden block
template <int S>
__device__ __forceinline__ Streams<S> layer_inputs(const GenParams& g, float* scr, int l) {
  Streams<S> r;
  if (l == 0) {
#pragma unroll
    for (int s = 0; s < S; ++s) r.p[s] = scr + g.sd.in0 + s * 8 * T;
  } else {
    const int npw = g.net.np[l];
    float* blk = scr + g.sd.hid[l - 1];
    r.p[0] = blk;  // a
#pragma unroll
    for (int s = 1; s < S; ++s) r.p[s] = blk + (S - 1 + s) * npw * T;  // H_x, H_t, H_xx
  }
  return r;
}

// packed fp32 FMA (fma.rn.f32x2 -> FFMA2, sm_100+): two IEEE fp32 FMAs per issue slot
__device__ __forceinline__ float2 ffma2(const float2 a, const float2 b, const float2 c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;"
      : "=l"(d)
      : "l"(*reinterpret_cast<const unsigned long long*>(&a)), "l"(*reinterpret_cast<const unsigned long long*>(&b)),
        "l"(*reinterpret_cast<const unsigned long long*>(&c)));
  return *reinterpret_cast<float2*>(&d);
}

__device__ __forceinline__ void cp_async16(float* dst, const float* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// rows [k0, k0+kc) of all S input streams -> act[s][i][32 points] (asynchronous; the caller waits)
template <int S>
__device__ __forceinline__ void stage_act(const Streams<S>& in, int k0, int kc, float* act, int kch) {
#pragma unroll
  for (int s = 0; s < S; ++s)
    for (int idx = threadIdx.x; idx < kc * 8; idx += blockDim.x) {
      const int i = idx >> 3, c = idx & 7;
      cp_async16(act + (s * kch + i) * T + c * 4, in.p[s] + (size_t)(k0 + i) * T + c * 4);
    }
}

// acc[s][jj] = sum_i in_s[i][lane] * W[i][8*jg + jj].  Operands are staged in shared memory with cp.async, K chunk by
// K chunk (one chunk unless the net is wider than the shared-memory budget): the input rows by the whole CTA, the
// 8-column weight slice of output group jg by the warp that owns it, so one L2 round trip feeds a whole chunk instead
// of one per unrolled iteration.  Called by every thread of the CTA (it contains CTA barriers); `active` warps compute.
// act_ready: the single chunk of inputs is already in flight / resident in `act` (later rounds, or pre-staged).
template <int S>
__device__ __forceinline__ void gemm_staged(const Streams<S>& in, int n_in, const float* __restrict__ W, int ldw, int jg,
                                            bool active, bool act_ready, float* act, float* wsl, int kch, int lane,
                                            float (&acc)[S][8]) {
  float2 a2[S][4];
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) a2[s][q] = make_float2(0.f, 0.f);
  const bool single = n_in <= kch;
  for (int k0 = 0; k0 < n_in; k0 += kch) {
    const int kc = min(kch, n_in - k0);
    if (!(single && act_ready)) stage_act<S>(in, k0, kc, act, kch);
    if (active) {
      const float* wsrc = W + (size_t)k0 * ldw + jg * 8;
      for (int idx = lane; idx < kc * 2; idx += 32) {
        const int row = idx >> 1, half = idx & 1;
        cp_async16(wsl + row * 8 + half * 4, wsrc + (size_t)row * ldw + half * 4);
      }
    }
    TRACE(200);
    cp_async_wait_all();
    TRACE(201);
    __syncthreads();
    TRACE(202);
    if (active) {
#pragma unroll 4
      for (int i = 0; i < kc; ++i) {
        const float4 w0 = *reinterpret_cast<const float4*>(wsl + i * 8);
        const float4 w1 = *reinterpret_cast<const float4*>(wsl + i * 8 + 4);
        const float2 wa = make_float2(w0.x, w0.y), wb = make_float2(w0.z, w0.w), wc = make_float2(w1.x, w1.y),
                     wd = make_float2(w1.z, w1.w);
#pragma unroll
        for (int s = 0; s < S; ++s) {
          const float x = act[(s * kch + i) * T + lane];
          const float2 xx = make_float2(x, x);
          a2[s][0] = ffma2(xx, wa, a2[s][0]);
          a2[s][1] = ffma2(xx, wb, a2[s][1]);
          a2[s][2] = ffma2(xx, wc, a2[s][2]);
          a2[s][3] = ffma2(xx, wd, a2[s][3]);
        }
      }
    }
    if (single) __syncwarp();  // the weight slice is warp private
    else __syncthreads();      // the next chunk overwrites act
  }
#pragma unroll
  for (int s = 0; s < S; ++s)
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      acc[s][2 * q] = a2[s][q].x;
      acc[s][2 * q + 1] = a2[s][q].y;
    }
}

// point-major ("transposed") copies of the weight-gradient operands: [stream][point][neuron], row stride ld.  The
// producers (forward / reverse epilogues) write them next to the neuron-major blocks the GEMMs read, so staging a
// weight-gradient operand is a plain 16-byte cp.async row copy instead of a transposing gather.
template <int S>
struct StreamsT {
  const float* p[S];
  int ld;
};

template <int S>
__device__ __forceinline__ StreamsT<S> layer_inputs_T(const GenParams& g, float* scr, int l) {
  StreamsT<S> r;
  if (l == 0) {
    r.ld = 8;
#pragma unroll
    for (int s = 0; s < S; ++s) r.p[s] = scr + g.sd.in0T + s * T * 8;
  } else {
    r.ld = g.net.np[l];
#pragma unroll
    for (int s = 0; s < S; ++s) r.p[s] = scr + g.sd.hidT[l - 1] + s * T * r.ld;
  }
  return r;
}

// 16-byte chunk c of a Z-bar row sits at chunk position zswz(c): the weight-gradient threads of a warp read the two chunks of
// their 8 columns, 2 jg and 2 jg + 1 for up to 25 consecutive jg; unswizzled those are the even (odd) positions only = 4 of
// the 8 bank groups, ~7 wavefronts per LDS.128; with bit 0 flipped in every other group of eight, 8 consecutive jg cover all
// 8 groups: 4 wavefronts, the minimum for 25 distinct chunks (the step was bound by the LSU, not by its FFMA2)
__device__ __forceinline__ int zswz(int c) { return c ^ ((c >> 3) & 1); }

// one stream's operands -> buf: Hs [T][np_in + 4] | Zs [T][np_out + 4] (chunks at zswz)   (asynchronous)
__device__ __forceinline__ void stage_wg(float* buf, const float* hsrc, int ld_h, int np_in, const float* zsrc, int ld_z,
                                         int np_out) {
  const int ldh = np_in + 4, ldz = np_out + 4;
  float* Hs = buf;
  float* Zs = buf + T * ldh;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int p = warp; p < T; p += nwarps) {
    for (int c = lane; c < np_in / 4; c += 32) cp_async16(Hs + p * ldh + c * 4, hsrc + (size_t)p * ld_h + c * 4);
    for (int c = lane; c < np_out / 4; c += 32) cp_async16(Zs + p * ldz + zswz(c) * 4, zsrc + (size_t)p * ld_z + c * 4);
  }
}

// W-bar_l += Hin^T Z-bar summed over the tile's points and all streams; b-bar_l += sum_p Z-bar_0.
// The per-stream operand tiles are double buffered when shared memory allows (nbuf == 2): stream s+1 lands while the
// 8x8 register tiles of stream s are accumulated.
template <int S>
__device__ void weight_grad(const GenParams& g, const StreamsT<S>& hin, const float* zbT, int l, float* gp, float* smem,
                            int nbuf, int crank, int cs, bool first) {
  const int n_in = g.net.n[l], n_out = g.net.n[l + 1];
  const int np_in = g.net.np[l], np_out = g.net.np[l + 1];
  const int ldh = np_in + 4, ldz = np_out + 4;
  const int bufsz = T * (2 * g.net.npmax + 8);
  const int ld_z = g.net.npmax;
  const int nti = np_in / 8, ntj = np_out / 8;
  float* gW = gp + g.net.w_off[l];
  float* gb = gp + g.net.b_off[l];
  const int ntasks = nti * ntj;
  const bool one_task = ntasks <= cs * (int)blockDim.x;  // then the 8x8 accumulator tile lives across the stream loop
  float2 acc[8][4];
  __syncthreads();  // earlier users of the staging region are done
  stage_wg(smem, hin.p[0], hin.ld, np_in, zbT, ld_z, np_out);
  for (int s = 0; s < S; ++s) {
    float* buf = smem + (nbuf == 2 ? (s & 1) * bufsz : 0);
    const float* Hs = buf;
    const float* Zs = buf + T * ldh;
    TRACE(210 + s);
    cp_async_wait_all();
    __syncthreads();  // stream s landed; every thread is past stream s-1, whose buffer may be refilled
    TRACE(220 + s);
    if (nbuf == 2 && s + 1 < S)
      stage_wg(smem + ((s + 1) & 1) * bufsz, hin.p[s + 1], hin.ld, np_in, zbT + (size_t)(s + 1) * T * ld_z, ld_z, np_out);
    // fewer 8x8 tiles than threads in the cluster (625 for 200 x 200 at the reference's batch sizes): every rank takes a
    // contiguous share instead of the first ranks taking a full block each and the rest idling
    const int share = one_task ? (ntasks + cs - 1) / cs : (int)blockDim.x;
    const int task0 = one_task ? (((int)threadIdx.x < share) ? crank * share + (int)threadIdx.x : ntasks)
                               : crank * (int)blockDim.x + (int)threadIdx.x;
    for (int task = task0; task < ntasks; task += cs * blockDim.x) {
      const int ig = task / ntj, jg = task % ntj;
      if (!one_task || s == 0) {
#pragma unroll
        for (int a = 0; a < 8; ++a)
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = make_float2(0.f, 0.f);
      }
#pragma unroll 4
      for (int p = 0; p < T; ++p) {
        const float4 h0 = *reinterpret_cast<const float4*>(Hs + p * ldh + ig * 8);
        const float4 h1 = *reinterpret_cast<const float4*>(Hs + p * ldh + ig * 8 + 4);
        const float4 z0 = *reinterpret_cast<const float4*>(Zs + p * ldz + zswz(2 * jg) * 4);
        const float4 z1 = *reinterpret_cast<const float4*>(Zs + p * ldz + zswz(2 * jg + 1) * 4);
        const float h[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const float2 z[4] = {make_float2(z0.x, z0.y), make_float2(z0.z, z0.w), make_float2(z1.x, z1.y), make_float2(z1.z, z1.w)};
#pragma unroll
        for (int a = 0; a < 8; ++a) {
          const float2 hh = make_float2(h[a], h[a]);
#pragma unroll
          for (int b = 0; b < 4; ++b) acc[a][b] = ffma2(hh, z[b], acc[a][b]);
        }
      }
      TRACE(230 + s);
      if (one_task && s < S - 1) continue;  // keep accumulating over the streams in registers
      // the partial row starts at zero: plain stores on the first visit, else one batch of 8 loads per row
      const bool plain = first && (one_task || s == 0);
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        const int i = ig * 8 + a;
        if (i < n_in) {
          float* row = gW + (size_t)i * n_out + jg * 8;
          float old[8];
#pragma unroll
          for (int b = 0; b < 8; ++b) old[b] = (!plain && jg * 8 + b < n_out) ? row[b] : 0.f;
#pragma unroll
          for (int b = 0; b < 8; ++b)
            if (jg * 8 + b < n_out) row[b] = old[b] + ((b & 1) ? acc[a][b / 2].y : acc[a][b / 2].x);
        }
      }
    }
    if (s == 0 && crank == 0) {
      for (int j = threadIdx.x; j < n_out; j += blockDim.x) {
        float sum = 0.f;
#pragma unroll 8
        for (int p = 0; p < T; ++p) sum += Zs[p * ldz + zswz(j >> 2) * 4 + (j & 3)];
        gb[j] = first ? sum : gb[j] + sum;
      }
    }
    if (nbuf == 1 && s + 1 < S) {
      __syncthreads();
      stage_wg(smem, hin.p[s + 1], hin.ld, np_in, zbT + (size_t)(s + 1) * T * ld_z, ld_z, np_out);
    }
  }
  __syncthreads();
}

struct PointSums {
  float v[PINN_NSUMS];
  float dl1, dl2;
};

// residual adjoint f_bar and loss bookkeeping for one residual value (appendix A.3)
__device__ __forceinline__ float seed_one(const GenParams& g, float f, float cB, bool valid, int64_t idx, PointSums& ps) {
  const LossCoef& lc = g.lc;
  float fbar = lc.cA * f;
  float zz = 0.f, gg = 0.f;
  const bool admm = (lc.loss == PINN_LOSS_V2_INF_ADMM || lc.loss == PINN_LOSS_V5_ADMM);
  if (admm && valid) {
    zz = g.z[idx];
    gg = g.gamma[idx];
    if (g.admm_op >= 4) {
      // z/gamma update of the previous epoch folded into this training pass (AB-ADMM:225-226 and :213 of the next
      // iteration, EUL:237-242 and :229, evaluate the same f): update first, seed and loss terms see the new state
      const float rho = lc.rho;
      const float kappa = 1.0f / (rho * (float)g.nf_global);
      if (g.admm_op == 5) gg = gg + rho * (f - zz);  // INF-ADMM:106-107: the dual advances inside z_update
      const float val = f + gg / rho;
      const float c1 = (val > kappa) ? 1.f : 0.f;
      const float c3 = (val < -1.0f * kappa) ? 1.f : 0.f;
      const float znew = c1 * (val - kappa) + c3 * (val + kappa);
      gg = gg + rho * (f - znew);
      zz = znew;
      g.z[idx] = zz;
      g.gamma[idx] = gg;
    }
  }
  const float sg = (f > 0.f) ? 1.f : ((f < 0.f) ? -1.f : 0.f);
  fbar += cB * sg + lc.cC * (f - zz) + lc.cD * gg;
  if (valid) {
    ps.v[PINN_SUM_F2] += f * f;
    ps.v[PINN_SUM_ABSF] += fabsf(f);
    if (admm) {
      const float t = f - zz + gg / lc.rho;
      float c = 0.5f * lc.rho * t * t;
      if (lc.loss == PINN_LOSS_V2_INF_ADMM) c += gg * f;
      ps.v[PINN_SUM_RES] += c;
      ps.v[PINN_SUM_MISFIT] += fabsf(f - zz);
    } else if (lc.loss == PINN_LOSS_V1_INF_L2 || lc.loss == PINN_LOSS_V4_MSE) {
      ps.v[PINN_SUM_RES] += f * f * lc.inv_nf;
    }
  } else {
    fbar = 0.f;
  }
  return fbar;
}

// soft threshold + dual update (AB-ADMM:185-198,:132; INF-ADMM:205-215,:106-107; EUL:203-215,:137-139)
__device__ __forceinline__ void admm_one(const GenParams& g, float f, int64_t idx) {
  if (g.admm_op >= 4) return;  // already applied inside seed_one
  if (g.admm_op == 1) {  // z <- f(theta0)  (AB-ADMM:96-97)
    g.z[idx] = f;
    return;
  }
  const float rho = g.lc.rho;
  const float kappa = 1.0f / (rho * (float)g.nf_global);
  float zz = g.z[idx], gg = g.gamma[idx];
  if (g.admm_op == 3) gg = gg + rho * (f - zz);  // INF-ADMM quirk: the dual advances inside z_update
  const float val = f + gg / rho;
  const float c1 = (val > kappa) ? 1.f : 0.f;
  const float c3 = (val < -1.0f * kappa) ? 1.f : 0.f;
  const float znew = c1 * (val - kappa) + c3 * (val + kappa);
  g.z[idx] = znew;
  g.gamma[idx] = gg + rho * (f - znew);
}

// the work of CTA `bid` of `nblk` on job g (a launch may carry two jobs: residual tiles and data-term tiles)
template <int S>
__device__ __forceinline__ void generic_body(const GenParams& g, int bid, int nblk, float* smem) {
  const NetDesc& net = g.net;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int cs = g.cluster;                       // CTAs per tile
  const int crank = bid % cs, cid = bid / cs, nclusters = nblk / cs;
  float* scr = g.scratch + (size_t)cid * g.sd.total;
  float* gp = g.part + (size_t)cid * g.rvlen;  // one partial row per cluster: its CTAs own disjoint entries
  const bool backward = (g.mode != GEN_MODE_FORWARD);
  const int L = net.L;
  // shared memory: act [S][kch][32] (GEMM inputs) | { per-warp weight slices [nwarps][kch][8]  or  the weight-gradient
  // staging tiles } (never live at the same time)
  const int kch = g.kch;
  float* act = smem;
  float* gst = smem + S * kch * T;
  float* wsl = gst + warp * kch * 8;
  const int per_round = cs * nwarps;
  // output group (8 neurons) of this warp in round r.  With fewer groups than warps in the cluster (25 for width 200 vs 64)
  // every rank takes a contiguous share -- one or two warps per scheduler on all eight SMs -- instead of the first ranks
  // filling all their warps and the rest idling
  auto group_of = [&](int r, int rank, int w, int ngroups) {
    if (ngroups <= per_round) {
      const int share = (ngroups + cs - 1) / cs;
      return (w < share) ? rank * share + w : ngroups;
    }
    return (r * cs + rank) * nwarps + w;
  };

  // a training pass overwrites every weight / bias entry on the CTA's first tile (plain stores in weight_grad), so
  // only the lambda gradients and the sums need a zero; a forward-only pass leaves the gradient part defined as zero
  if (backward) {
    if (crank == 0)
      for (int k = net.P + threadIdx.x; k < g.rvlen; k += blockDim.x) gp[k] = 0.f;
  } else {
    for (int k = crank * blockDim.x + threadIdx.x; k < g.rvlen; k += cs * blockDim.x) gp[k] = 0.f;
  }
  float cB = g.lc.cB;
  if (g.lc.loss == PINN_LOSS_V3_L1SQ && g.l1_sum != nullptr) cB = 2.0f * g.lc.inv_nf * g.l1_sum[0];
  const float lam1 = g.theta[net.P], lam2 = g.theta[net.P + 1];
  PointSums ps;
#pragma unroll
  for (int k = 0; k < PINN_NSUMS; ++k) ps.v[k] = 0.f;
  ps.dl1 = ps.dl2 = 0.f;
  __syncthreads();

  const int64_t ntiles = (g.N + T - 1) / T;
  bool first = true;
  for (int64_t tile = cid; tile < ntiles; tile += nclusters) {
    const int64_t pidx = tile * T + lane;
    const bool valid = pidx < g.N;
    // ---- Taylor seeds of the input layer (appendix A.2): H0 = 2(X-lb)/(ub-lb)-1 ----
    if (warp == 0 && crank == 0) {
      float x = net.lbx, t = net.lbt;
      if (valid) {
        const float2 xt = *reinterpret_cast<const float2*>(g.X + 2 * pidx);
        x = xt.x;
        t = xt.y;
      }
      float* in0 = scr + g.sd.in0;
      in0[0 * T + lane] = 2.0f * (x - net.lbx) / net.spanx - 1.0f;
      in0[1 * T + lane] = 2.0f * (t - net.lbt) / net.spant - 1.0f;
      if (S >= 3) {
        in0[(1 * 8 + 0) * T + lane] = 2.0f / net.spanx;
        in0[(1 * 8 + 1) * T + lane] = 0.f;
        in0[(2 * 8 + 0) * T + lane] = 0.f;
        in0[(2 * 8 + 1) * T + lane] = 2.0f / net.spant;
      }
      if (S == 4) {
        in0[(3 * 8 + 0) * T + lane] = 0.f;
        in0[(3 * 8 + 1) * T + lane] = 0.f;
      }
      if (backward) {  // point-major copy for the layer-0 weight gradient
        float4* in0T = reinterpret_cast<float4*>(scr + g.sd.in0T);
        const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
        in0T[(0 * T + lane) * 2] = make_float4(in0[0 * T + lane], in0[1 * T + lane], 0.f, 0.f);
        in0T[(0 * T + lane) * 2 + 1] = zero4;
        if (S >= 3) {
          in0T[(1 * T + lane) * 2] = make_float4(2.0f / net.spanx, 0.f, 0.f, 0.f);
          in0T[(1 * T + lane) * 2 + 1] = zero4;
          in0T[(2 * T + lane) * 2] = make_float4(0.f, 2.0f / net.spant, 0.f, 0.f);
          in0T[(2 * T + lane) * 2 + 1] = zero4;
        }
        if (S == 4) {
          in0T[(3 * T + lane) * 2] = zero4;
          in0T[(3 * T + lane) * 2 + 1] = zero4;
        }
      }
    }
    tile_sync(cs);
    TRACE(1);

    // ---- forward ----
    for (int l = 0; l < L; ++l) {
      const Streams<S> in = layer_inputs<S>(g, scr, l);
      const int n_in = net.n[l], np_out = net.np[l + 1], n_out = net.n[l + 1];
      const float* W = g.wp + net.wp_off[l];
      const float* b = g.theta + net.b_off[l];
      const bool head = (l == L - 1);
      const int ngroups = np_out / 8;
      for (int r = 0; r * per_round < ngroups; ++r) {
        const int jg = group_of(r, crank, warp, ngroups);
        const bool active = jg < ngroups;
        float acc[S][8];
        TRACE(100 + l);
        gemm_staged<S>(in, n_in, W, np_out, jg, active, r > 0, act, wsl, kch, lane, acc);
        TRACE(120 + l);
        if (!active) continue;
        if (!head) {
          float* blk = scr + g.sd.hid[l];
          float hT[S][8];  // the layer's output streams a, H_x, H_t, H_xx of this point
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int j = jg * 8 + jj;
            const float z = acc[0][jj] + (j < n_out ? b[j] : 0.f);
            const float a = pinn_tanh(z);
            const float d1 = fmaf(-a, a, 1.0f);
            blk[j * T + lane] = a;
            hT[0][jj] = a;
            if (S >= 3) {
              const float zx = acc[1][jj], zt = acc[2][jj];
              blk[(1 * np_out + j) * T + lane] = zx;
              blk[(2 * np_out + j) * T + lane] = zt;
              blk[((S - 1 + 1) * np_out + j) * T + lane] = hT[1][jj] = d1 * zx;
              blk[((S - 1 + 2) * np_out + j) * T + lane] = hT[2][jj] = d1 * zt;
              if (S == 4) {
                const float zxx = acc[3][jj];
                blk[(3 * np_out + j) * T + lane] = zxx;
                blk[((S - 1 + 3) * np_out + j) * T + lane] = hT[3][jj] = d1 * fmaf(-2.0f * a, zx * zx, zxx);
              }
            }
          }
          if (backward) {
            float* hidT = scr + g.sd.hidT[l];
#pragma unroll
            for (int s = 0; s < S; ++s) {
              float4* dst = reinterpret_cast<float4*>(hidT + ((size_t)s * T + lane) * np_out + jg * 8);
              dst[0] = make_float4(hT[s][0], hT[s][1], hT[s][2], hT[s][3]);
              dst[1] = make_float4(hT[s][4], hT[s][5], hT[s][6], hT[s][7]);
            }
          }
        } else {
          float* Y = scr + g.sd.Y;
#pragma unroll
          for (int jj = 0; jj < 8; ++jj) {
            const int j = jg * 8 + jj;  // jg == 0: n_out <= 8
            Y[j * T + lane] = acc[0][jj] + (j < n_out ? b[j] : 0.f);
#pragma unroll
            for (int s = 1; s < S; ++s) Y[(s * 8 + j) * T + lane] = acc[s][jj];
          }
        }
      }
      TRACE(140 + l);
      tile_sync(cs);
      TRACE(10 + l);
    }

    // ---- residual, loss terms, ADMM, adjoint seeds of the head outputs ----
    int cur = 0;
    if (warp == 0 && crank == 0) {
      const float* Y = scr + g.sd.Y;
      float* zb = scr + g.sd.zb[0];
      float* zbT = scr + g.sd.zbT[0];  // [s][point][npmax]
      const int ldT = net.npmax;
      const int ldz = net.npmax * T;
      if (S == 1) {  // data term: adjoints supplied by the caller (appendix A.3, dL/du^)
        for (int o = 0; o < 8; ++o) {
          float sd = 0.f;
          if (o < net.n_out && valid) {
            if (g.u_out) g.u_out[pidx * net.n_out + o] = Y[o * T + lane];
            if (g.seed) sd = g.seed[pidx * net.n_out + o];
            if (g.u_data) {  // squared misfit data_c * sum (u^ - u)^2 in one pass (all losses but INF-L2's norm)
              const float r = Y[o * T + lane] - g.u_data[pidx * net.n_out + o];
              sd = 2.0f * g.data_c * r;
              ps.v[PINN_SUM_DATA] += g.data_c * r * r;
            }
          }
          zb[o * T + lane] = sd;
          zbT[(size_t)lane * ldT + o] = sd;
        }
      } else if (net.pde == PINN_PDE_BURGERS) {
        const float u = Y[lane], ux = Y[(1 * 8) * T + lane], ut = Y[(2 * 8) * T + lane];
        const float uxx = (S == 4) ? Y[(3 * 8) * T + lane] : 0.f;
        const float f = ut + lam1 * u * ux - lam2 * uxx;  // INF-L2:118 / AB-ADMM:178
        if (valid) {
          if (g.u_out) g.u_out[pidx] = u;
          if (g.f_out) g.f_out[pidx] = f;
        }
        const float fbar = seed_one(g, f, cB, valid, pidx, ps);
        if (valid && g.admm_op) admm_one(g, f, pidx);
        ps.dl1 += fbar * u * ux;
        ps.dl2 -= fbar * uxx;
        for (int o = 0; o < 8; ++o) {
          const bool o0 = (o == 0);
          const float v0 = o0 ? fbar * lam1 * ux : 0.f, v1 = o0 ? fbar * lam1 * u : 0.f, v2 = o0 ? fbar : 0.f;
          const float v3 = o0 ? -lam2 * fbar : 0.f;
          zb[0 * ldz + o * T + lane] = v0;
          zb[1 * ldz + o * T + lane] = v1;
          zb[2 * ldz + o * T + lane] = v2;
          zbT[((size_t)0 * T + lane) * ldT + o] = v0;
          zbT[((size_t)1 * T + lane) * ldT + o] = v1;
          zbT[((size_t)2 * T + lane) * ldT + o] = v2;
          if (S == 4) {
            zb[3 * ldz + o * T + lane] = v3;
            zbT[((size_t)3 * T + lane) * ldT + o] = v3;
          }
        }
      } else {  // Euler, outputs (rho,u,E), EUL:176-198 by the product rule
        const float k = 0.4f;
        const float r = Y[0 * T + lane], u = Y[1 * T + lane], E = Y[2 * T + lane];
        const float rx = Y[(8 + 0) * T + lane], ux = Y[(8 + 1) * T + lane], Ex = Y[(8 + 2) * T + lane];
        const float rt = Y[(16 + 0) * T + lane], ut = Y[(16 + 1) * T + lane], Et = Y[(16 + 2) * T + lane];
        const float p = k * (E - 0.5f * r * u * u);
        const float px = k * (Ex - 0.5f * rx * u * u - r * u * ux);
        const float f1 = rt + rx * u + r * ux;
        const float f2 = rt * u + r * ut + rx * u * u + 2.0f * r * u * ux + px;
        const float f3 = Et + ux * E + u * Ex + ux * p + u * px;
        if (valid) {
          if (g.u_out) {
            g.u_out[pidx * 3 + 0] = r;
            g.u_out[pidx * 3 + 1] = u;
            g.u_out[pidx * 3 + 2] = E;
          }
          if (g.f_out) {
            g.f_out[pidx * 3 + 0] = f1;
            g.f_out[pidx * 3 + 1] = f2;
            g.f_out[pidx * 3 + 2] = f3;
          }
        }
        const float b1 = seed_one(g, f1, cB, valid, pidx * 3 + 0, ps);
        const float b2 = seed_one(g, f2, cB, valid, pidx * 3 + 1, ps);
        const float b3 = seed_one(g, f3, cB, valid, pidx * 3 + 2, ps);
        if (valid && g.admm_op) {
          admm_one(g, f1, pidx * 3 + 0);
          admm_one(g, f2, pidx * 3 + 1);
          admm_one(g, f3, pidx * 3 + 2);
        }
        const float p_r = -0.5f * k * u * u, p_u = -k * r * u, p_E = k;
        const float px_r = -k * u * ux, px_u = -k * (rx * u + r * ux);
        const float px_rx = -0.5f * k * u * u, px_ux = -k * r * u, px_Ex = k;
        float yb[3][3];
        yb[0][0] = b1 * ux + b2 * (ut + 2.0f * u * ux + px_r) + b3 * (ux * p_r + u * px_r);
        yb[0][1] = b1 * rx + b2 * (rt + 2.0f * rx * u + 2.0f * r * ux + px_u) + b3 * (Ex + ux * p_u + px + u * px_u);
        yb[0][2] = b3 * (ux + ux * p_E);
        yb[1][0] = b1 * u + b2 * (u * u + px_rx) + b3 * u * px_rx;
        yb[1][1] = b1 * r + b2 * (2.0f * r * u + px_ux) + b3 * (E + p + u * px_ux);
        yb[1][2] = b2 * px_Ex + b3 * (u + u * px_Ex);
        yb[2][0] = b1 + b2 * u;
        yb[2][1] = b2 * r;
        yb[2][2] = b3;
        for (int s = 0; s < 3; ++s)
          for (int o = 0; o < 8; ++o) {
            const float v = (o < 3) ? yb[s][o] : 0.f;
            zb[s * ldz + o * T + lane] = v;
            zbT[((size_t)s * T + lane) * ldT + o] = v;
          }
      }
    }
    tile_sync(cs);
    TRACE(30);
    if (!backward) continue;

    // ---- reverse sweep ----
    for (int l = L - 1; l >= 0; --l) {
      const float* zb = scr + g.sd.zb[cur];
      const StreamsT<S> hinT = layer_inputs_T<S>(g, scr, l);
      Streams<S> zin;
#pragma unroll
      for (int s = 0; s < S; ++s) zin.p[s] = zb + (size_t)s * net.npmax * T;
      const bool prestage = (l > 0) && (net.n[l + 1] <= kch);
      if (prestage) stage_act<S>(zin, 0, net.n[l + 1], act, kch);  // lands while the weight gradient runs
      weight_grad<S>(g, hinT, scr + g.sd.zbT[cur], l, gp, gst, g.wg_nbuf, crank, cs, first);
      TRACE(40 + l);
      if (l > 0) {
        const int n_j = net.n[l + 1], np_i = net.np[l];
        const float* WT = g.wt + net.wt_off[l];
        const float* blk = scr + g.sd.hid[l - 1];
        float* zn = scr + g.sd.zb[cur ^ 1];
        float* znT = scr + g.sd.zbT[cur ^ 1];
        const int ldz = net.npmax * T;
        const int ngroups = np_i / 8;
        for (int r = 0; r * per_round < ngroups; ++r) {
          const int ig = group_of(r, crank, warp, ngroups);
          const bool active = ig < ngroups;
          float acc[S][8];
          gemm_staged<S>(zin, n_j, WT, np_i, ig, active, prestage || r > 0, act, wsl, kch, lane, acc);
          if (!active) continue;
          float zT[S][8];  // this point's new adjoints, for the point-major copy
#pragma unroll
          for (int ii = 0; ii < 8; ++ii) {
            const int i = ig * 8 + ii;
            const float a = blk[i * T + lane];
            const float d1 = fmaf(-a, a, 1.0f);
            const float hb = acc[0][ii];
            if (S == 1) {
              zn[i * T + lane] = zT[0][ii] = d1 * hb;
            } else {
              const float d2 = -2.0f * a * d1;
              const float zx = blk[(1 * np_i + i) * T + lane], zt = blk[(2 * np_i + i) * T + lane];
              const float hxb = acc[1][ii], htb = acc[2][ii];
              float zbar = d1 * hb + d2 * (zx * hxb + zt * htb);
              float zxbar = d1 * hxb;
              const float ztbar = d1 * htb;
              if (S == 4) {
                const float zxx = blk[(3 * np_i + i) * T + lane];
                const float hxxb = acc[3][ii];
                const float d3 = -2.0f * d1 * fmaf(-3.0f * a, a, 1.0f);
                zbar += d2 * zxx * hxxb + d3 * zx * zx * hxxb;
                zxbar += 2.0f * d2 * zx * hxxb;
                zn[3 * ldz + i * T + lane] = zT[S - 1][ii] = d1 * hxxb;
              }
              zn[0 * ldz + i * T + lane] = zT[0][ii] = zbar;
              zn[1 * ldz + i * T + lane] = zT[S > 1 ? 1 : 0][ii] = zxbar;
              zn[2 * ldz + i * T + lane] = zT[S > 2 ? 2 : 0][ii] = ztbar;
            }
          }
#pragma unroll
          for (int s = 0; s < S; ++s) {
            float4* dst = reinterpret_cast<float4*>(znT + ((size_t)s * T + lane) * net.npmax + ig * 8);
            dst[0] = make_float4(zT[s][0], zT[s][1], zT[s][2], zT[s][3]);
            dst[1] = make_float4(zT[s][4], zT[s][5], zT[s][6], zT[s][7]);
          }
        }
        tile_sync(cs);
        TRACE(60 + l);
        cur ^= 1;
      }
    }
    cur = 0;
    first = false;
  }

  if (first && backward) {  // no tile reached this CTA (the host never sizes a grid that way): keep the row defined
    for (int k = crank * blockDim.x + threadIdx.x; k < net.P; k += cs * blockDim.x) gp[k] = 0.f;
  }
  // ---- per-cluster partial sums (the residual / data bookkeeping is done by rank 0's first warp) ----
  if (warp == 0 && crank == 0) {
#pragma unroll
    for (int k = 0; k < PINN_NSUMS; ++k) {
      const float v = warp_sum(ps.v[k]);
      if (lane == 0) gp[net.P + 2 + k] += v;
    }
    const float d1 = warp_sum(ps.dl1), d2 = warp_sum(ps.dl2);
    if (lane == 0) {
      gp[net.P] += d1;
      gp[net.P + 1] += d2;
    }
  }
}

template <int S, int NT>
__global__ void __launch_bounds__(NT, NT > 256 ? 1 : 2) pinn_generic_kernel(const __grid_constant__ GenParams g) {
  extern __shared__ float smem[];
  generic_body<S>(g, blockIdx.x, gridDim.x, smem);
}

// residual tiles (CTAs [0, grid_res)) and data-term tiles (the rest) in one launch: at the reference's batch sizes
// neither job fills the GPU, and one launch + one partial-sum reduction replaces two of each
template <int S, int NT>
__global__ void __launch_bounds__(NT, NT > 256 ? 1 : 2)
    pinn_generic_dual_kernel(const __grid_constant__ GenParams g, const __grid_constant__ GenParams gd, int grid_res) {
  extern __shared__ float smem[];
  if ((int)blockIdx.x < grid_res) generic_body<S>(g, blockIdx.x, grid_res, smem);
  else generic_body<1>(gd, blockIdx.x - grid_res, gridDim.x - grid_res, smem);
}

}  // namespace

#ifdef PINN_TRACE
extern "C" int pinn_debug_trace(long long* out, int* n) {
  cudaDeviceSynchronize();
  cudaMemcpyFromSymbol(n, g_trace_n, sizeof(int));
  cudaMemcpyFromSymbol(out, g_trace, sizeof(long long) * 2048);
  int zero = 0;
  cudaMemcpyToSymbol(g_trace_n, &zero, sizeof(int));
  return 0;
}
#endif

// shared memory of one CTA: act [S][kch][T] + max(weight slices [warps][kch][8], nbuf weight-gradient staging buffers
// [T][2 npmax + 8]).  kch covers the widest layer unless that would pass ~216 KB; the staging is double buffered when it
// fits and (small grids) occupancy is not at stake
// 288 threads iff the ninth warp lowers the number of GEMM rounds of the widest layer for some cluster size the launcher uses
// (see GEN_THREADS_WIDE) and the net is too wide for two CTAs per SM anyway.  A property of the NET, not of the launch: the
// forward-only passes (z / gamma updates, predict) and the training pass of a handle must run the same instantiation -- the
// folded ADMM update is bit-identical to the two-pass form only if both evaluate the residuals with the same machine code.
size_t pinn_generic_smem_bytes(const NetDesc& net, int S, bool want_occupancy, int* kch_out, int* nbuf_out, int threads);
static int gen_threads_for(const NetDesc& net, int S) {
  int groups = 1;
  for (int l = 1; l <= net.L; ++l) groups = net.np[l] / 8 > groups ? net.np[l] / 8 : groups;
  bool fewer = false;
  for (int cs = 1; cs <= 4; ++cs) {
    auto rounds = [&](int warps) { return (groups + cs * warps - 1) / (cs * warps); };
    fewer = fewer || rounds(GEN_THREADS_WIDE / 32) < rounds(GEN_THREADS / 32);
  }
  const bool one_cta_per_sm = pinn_generic_smem_bytes(net, S > 1 ? S : 3, true, nullptr, nullptr, GEN_THREADS) > 113 * 1024;
  return (fewer && one_cta_per_sm) ? GEN_THREADS_WIDE : GEN_THREADS;
}

size_t pinn_generic_smem_bytes(const NetDesc& net, int S, bool want_occupancy, int* kch_out, int* nbuf_out, int threads) {
  const size_t gsz = (size_t)T * (2 * net.npmax + 8);
  const size_t budget = 216 * 1024 / sizeof(float);
  int kch = net.npmax;
  auto total = [&](int k, int nbuf) {
    const size_t w = (size_t)(threads / 32) * k * 8;
    return (size_t)S * k * T + (w > nbuf * gsz ? w : nbuf * gsz);
  };
  while (kch > 8 && total(kch, 1) > budget) kch -= 8;
  int nbuf = 1;
  const size_t half = 113 * 1024 / sizeof(float);  // two CTAs per SM below this
  if (kch == net.npmax && total(kch, 2) <= budget && !(want_occupancy && total(kch, 1) <= half && total(kch, 2) > half)) nbuf = 2;
  if (kch_out) *kch_out = kch;
  if (nbuf_out) *nbuf_out = nbuf;
  return total(kch, nbuf) * sizeof(float);
}

cudaError_t pinn_generic_launch(const GenParams& g_in, int S, int grid, cudaStream_t stream) {
  GenParams g = g_in;
  const int nt = gen_threads_for(g.net, S);
  const size_t smem = pinn_generic_smem_bytes(g.net, S, grid > 148, &g.kch, &g.wg_nbuf, nt);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(nt);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
#define LAUNCH(SS, NT)                                                                                              \
  e = cudaFuncSetAttribute(pinn_generic_kernel<SS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
  if (e != cudaSuccess) return e;                                                                                   \
  e = cudaLaunchKernelEx(&cfg, pinn_generic_kernel<SS, NT>, g);                                                     \
  if (e != cudaSuccess) return e;
  if (nt == GEN_THREADS) {
    if (S == 1) {
      LAUNCH(1, GEN_THREADS)
    } else if (S == 3) {
      LAUNCH(3, GEN_THREADS)
    } else {
      LAUNCH(4, GEN_THREADS)
    }
  } else {
    if (S == 1) {
      LAUNCH(1, GEN_THREADS_WIDE)
    } else if (S == 3) {
      LAUNCH(3, GEN_THREADS_WIDE)
    } else {
      LAUNCH(4, GEN_THREADS_WIDE)
    }
  }
#undef LAUNCH
  return cudaGetLastError();
}

// g: residual job on grid_res CTAs, gd: data-term job (S = 1) on grid_data CTAs; both use clusters of g.cluster CTAs
cudaError_t pinn_generic_dual_launch(const GenParams& g_in, int S, int grid_res, const GenParams& gd_in, int grid_data,
                                     cudaStream_t stream) {
  GenParams g = g_in, gd = gd_in;
  const int nt = gen_threads_for(g.net, S);
  const size_t smem_res = pinn_generic_smem_bytes(g.net, S, grid_res + grid_data > 148, &g.kch, &g.wg_nbuf, nt);
  const size_t smem_data = pinn_generic_smem_bytes(gd.net, 1, true, &gd.kch, &gd.wg_nbuf, nt);
  const size_t smem = smem_res > smem_data ? smem_res : smem_data;
  gd.cluster = g.cluster;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid_res + grid_data);
  cfg.blockDim = dim3(nt);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = g.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e;
#define LAUNCH(SS, NT)                                                                                                  \
  e = cudaFuncSetAttribute(pinn_generic_dual_kernel<SS, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);   \
  if (e != cudaSuccess) return e;                                                                                       \
  e = cudaLaunchKernelEx(&cfg, pinn_generic_dual_kernel<SS, NT>, g, gd, grid_res);                                      \
  if (e != cudaSuccess) return e;
  if (nt == GEN_THREADS) {
    if (S == 3) {
      LAUNCH(3, GEN_THREADS)
    } else {
      LAUNCH(4, GEN_THREADS)
    }
  } else {
    if (S == 3) {
      LAUNCH(3, GEN_THREADS_WIDE)
    } else {
      LAUNCH(4, GEN_THREADS_WIDE)
    }
  }
#undef LAUNCH
  return cudaGetLastError();
}

// How many clusters of `cs` CTAs the GPU holds with ONE CTA per SM (a cluster lives inside one GPC, so this is the
// sum over GPCs of floor(SMs / cs), e.g. ~34 for cs = 4 on 148 SMs, not 37): asked of the occupancy calculator with the
// dynamic shared memory set so high that a second CTA cannot join an SM.  0 if the query fails.
int pinn_generic_cluster_capacity(int cs) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cs * 64);
  cfg.blockDim = dim3(GEN_THREADS);
  cfg.dynamicSmemBytes = 120 * 1024;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (cudaFuncSetAttribute(pinn_generic_kernel<3, GEN_THREADS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 120 * 1024) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, pinn_generic_kernel<3, GEN_THREADS>, &cfg) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

