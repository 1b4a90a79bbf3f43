// kernels_feat_tc.cu -- visual (ReID feature) cost matrix: tensor-core screen (tcgen05.mma + TMEM + TMA) followed by
// an exact f32 refinement of the surviving pairs.
//
// What the reference computes (src/distance.rs:9-47, src/trackers/visual_sort/metric.rs:200-295): for every
// (candidate, track-observation) pair the f32 euclidean / cosine distance, kept only if it passes the metric's
// threshold (euclid d <= thr, cosine cos >= thr).  After thresholding the matrix is sparse: a detection is close to a
// handful of observations of "its" track and far from everything else.
//
// GPU shape:
//   1. screen  : C~[m][c] = sum_d A[m][d] * B[c][d] with BF16 operand copies on the 5th-gen tensor cores
//                (one tcgen05.mma.kind::f16 chain, fp32 accumulation in TMEM).  The BF16 rounding error of the dot
//                product is bounded by E = 2^-8 * ||a|| * ||b|| (Cauchy-Schwarz), so a pair can only pass the
//                threshold if its screened value is within E of it.  Everything else is written as None (NaN)
//                straight from the epilogue; survivors are appended to a compact pair list.
//   2. refine  : one warp per surviving pair recomputes the distance in f32 in the reference's exact summation order
//                (8-lane blocks, horizontal reduce_add, sequential block accumulation) and applies is_ok /
//                distance_to_weight.  Every value that reaches the voting stage is therefore bit-identical to the
//                CPU reference -- the tensor cores only decide which pairs are worth computing.
//   If the pair list overflows (a non-selective threshold) the caller falls back to the dense exact SIMT kernel.
//
// Screen kernel: persistent CTAs (one per SM), 192 threads = TMA producer warp, MMA issuer warp, 4 epilogue warps.
// Tile 128 (candidates) x 256 (track-observation rows) x 64 (features = one 128-byte swizzle atom of bf16);
// 4 smem stages of 48 KB; the 512 TMEM columns hold two 128x256 fp32 accumulators so the epilogue of tile i overlaps
// the MMAs of tile i+1.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "sb_engine.cuh"
#include "sb_tc.cuh"

namespace sb {

constexpr int TC_STAGES = 4;          // 48 KB stages (A tile + whole B tile)
constexpr int TC_STAGES_PAIR = 6;     // 32 KB stages of the cta_group::2 variant (A tile + half of the B tile)
constexpr int TC_STAGE_BYTES_PAIR = 32768;
constexpr int TC_STAGE_BYTES = TC_A_BYTES + TC_B_BYTES;  // 48 KB
constexpr int TC_THREADS = 320;  // TMA warp, MMA warp, 2 x 4 epilogue warps (the two groups alternate tiles)

struct TcHdr { int scene, m0, ncols_left, m, det_base, col0, epoch, vis_lbase, vis_lcap, pad; };
struct TcSmem {
  unsigned char stage[TC_STAGES][TC_STAGE_BYTES];  // 1024-byte aligned operand stages first
  // per epilogue group: the metadata slabs of its current tile, bulk-copied by the producer warp
  // four slab sets (two per group): the producer runs up to two tile pairs ahead of the epilogue
  VisColMeta meta[4][TC_BN];
  VisRowMeta rowm[4][TC_BM];
  float colb[4][TC_BN];
  unsigned int colvalid[4][TC_BN / 32];
  TcHdr hdr[4];
  unsigned long long full_bar[TC_STAGES_PAIR];
  unsigned long long empty_bar[TC_STAGES_PAIR];
  unsigned long long tmem_full[2];
  unsigned long long tmem_empty[2];
  unsigned long long meta_full[4];
  unsigned long long meta_empty[4];
  unsigned int tmem_base;
};


// ------------------------------------------------------------------------------------------------ screen kernel
// CL == 2: clusters of two CTAs work on two candidate tiles (m0, m0 + 128) of the same track-row tile; each CTA loads
// its own A tile and HALF of the B tile, multicast into both CTAs' shared memory, so B crosses L2 -> SM once per pair.
// CL == 3: the two CTAs form a cta_group::2 pair: one 256 x 256 x 16 MMA per instruction, issued by the leader CTA only.
// Each CTA stages its own A tile and HALF of the B tile (no multicast: the tensor cores read the peer's half), which
// halves the shared-memory bytes written and read per MMA -- with 128 x 256 single-CTA tiles every operand byte is
// written once by TMA and read once by the MMA, 192 B/clk against the 128 B/clk an SM's shared memory delivers.
template <int CL, bool COSINE>
__global__ void __launch_bounds__(TC_THREADS, 1)
vis_screen_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB, Params p,
                  TrackStore ts, Frame f, const TcTile* tiles, int n_tiles_host, const int* n_tiles_dev,
                  const VisColMeta* colmeta, const VisColGeo* colgeo, const VisRowMeta* rowmeta, const float* colb,
                  const unsigned int* colvalid) {
  // trackers: the tile list is built on the device (frame_setup_kernel) and so is its length
  const int n_tiles = n_tiles_dev ? *n_tiles_dev : n_tiles_host;
  extern __shared__ unsigned char smem_raw_[];
  // offset arithmetic on the __shared__ array (not on an integer) keeps the accesses in the shared address space
  TcSmem& S = *reinterpret_cast<TcSmem*>(smem_raw_ + ((1024u - (smem_u32(smem_raw_) & 1023u)) & 1023u));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = p.max_obs;
  const int KB = (p.d8 + TC_BK - 1) / TC_BK;

  constexpr bool PAIR = CL == 3;
  constexpr int NST = PAIR ? TC_STAGES_PAIR : TC_STAGES;
  constexpr int STAGE_B = PAIR ? TC_STAGE_BYTES_PAIR : TC_STAGE_BYTES;
  unsigned char* const stage_base = &S.stage[0][0];   // NST stages of STAGE_B bytes (192 KB either way)
  const uint32_t crank = CL >= 2 ? cluster_rank() : 0u;
  const int cta_first = CL >= 2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // first (cluster) tile of this CTA
  const int cta_step = CL >= 2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  if (threadIdx.x == 32) {
    for (int s = 0; s < NST; ++s) { mbar_init(&S.full_bar[s], 1); mbar_init(&S.empty_bar[s], CL == 2 ? 2 : 1); }
    for (int b = 0; b < 2; ++b) {
      // PAIR: the leader's tmem_empty collects the epilogue warps of both CTAs
      mbar_init(&S.tmem_full[b], 1); mbar_init(&S.tmem_empty[b], PAIR ? 8 : 4);
    }
    for (int b = 0; b < 4; ++b) { mbar_init(&S.meta_full[b], 1); mbar_init(&S.meta_empty[b], 4); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (PAIR) {   // both CTAs of the pair allocate, same warp, same destination
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::);
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&S.tmem_base)), "r"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();   // the peer's barriers are initialised before anything is multicast into them
  tc_fence_after();
  const uint32_t tmem_base = S.tmem_base;

  if (warp == 0) {
    // ===================================================================== TMA producer
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapA) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)&mapB) : "memory");
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      TcTile tl_n;
      tl_n.scene = 0; tl_n.m0 = 0; tl_n.c0 = 0; tl_n.pad = 0;
      if (cta_first < n_tiles) tl_n = tiles[cta_first];
      SceneDesc sc_n = f.scenes[tl_n.scene];
      for (int t = cta_first; t < n_tiles; t += cta_step, ++it) {
        const TcTile tl = tl_n;
        const SceneDesc sc = sc_n;
        if (t + cta_step < n_tiles) { tl_n = tiles[t + cta_step]; sc_n = f.scenes[tl_n.scene]; }
        const int m0 = tl.m0 + (int)crank * TC_BM;
        const int rowA = sc.det_base + m0;
        const int rowB = sc.slot * ts.track_cap * K + tl.c0;
        {
          // metadata of this tile for the epilogue group that will drain it: header by plain stores (published by the
          // release of the arrive below), column / row slabs by bulk copies that complete on the same barrier
          const int g = it & 3;   // slab set: tile parity picks the group, bit 1 alternates the group's two sets
          mbar_wait(&S.meta_empty[g], ((it >> 2) & 1) ^ 1);
          TcHdr h;
          h.scene = tl.scene; h.m0 = m0; h.ncols_left = sc.nb * K - tl.c0; h.m = sc.m; h.det_base = sc.det_base;
          h.col0 = sc.col_off + tl.c0; h.epoch = (int)sc.epoch; h.vis_lbase = sc.vis_lbase; h.vis_lcap = sc.vis_lcap; h.pad = 0;
          S.hdr[g] = h;
          mbar_expect_tx(&S.meta_full[g], (uint32_t)(sizeof(VisColMeta) * TC_BN + sizeof(VisRowMeta) * TC_BM +
                                                    4 * TC_BN + TC_BN / 8));
          bulk_load(S.meta[g], colmeta + h.col0, (uint32_t)(sizeof(VisColMeta) * TC_BN), &S.meta_full[g]);
          bulk_load(S.rowm[g], rowmeta + rowA, (uint32_t)(sizeof(VisRowMeta) * TC_BM), &S.meta_full[g]);
          bulk_load(S.colb[g], colb + h.col0, 4 * TC_BN, &S.meta_full[g]);
          bulk_load(S.colvalid[g], colvalid + (h.col0 >> 5), TC_BN / 8, &S.meta_full[g]);   // col0 is a multiple of 128
        }
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&S.empty_bar[stage], phase ^ 1);   // CL == 2: both CTAs have released the stage
          unsigned char* base = stage_base + stage * STAGE_B;
          if (PAIR) {
            // both CTAs' bytes are counted on the leader's barrier, which the leader arms for the whole pair
            const uint32_t lbar = leader_addr(&S.full_bar[stage]);
            if (crank == 0) mbar_expect_tx(&S.full_bar[stage], 2 * TC_STAGE_BYTES_PAIR);
            tma_load_2d_pair(base, &mapA, kb * TC_BK, rowA, lbar);
            tma_load_2d_pair(base + TC_A_BYTES, &mapB, kb * TC_BK, rowB + (int)crank * (TC_BN / 2), lbar);
          } else {
            mbar_expect_tx(&S.full_bar[stage], TC_STAGE_BYTES);
            tma_load_2d(base, &mapA, kb * TC_BK, rowA, &S.full_bar[stage]);
            if (CL == 2) {
              // this CTA's half of the B tile (rows rank*128 .. +128), delivered to both CTAs
              tma_load_2d_mc(base + TC_A_BYTES + crank * (TC_B_BYTES / 2), &mapB, kb * TC_BK, rowB + (int)crank * (TC_BN / 2),
                             &S.full_bar[stage], (uint16_t)0x3);
            } else {
              tma_load_2d(base + TC_A_BYTES, &mapB, kb * TC_BK, rowB, &S.full_bar[stage]);
            }
          }
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================================================================== MMA issuer (one elected lane; PAIR: leader CTA)
    if (lane == 0 && (!PAIR || crank == 0)) {
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int t = cta_first; t < n_tiles; t += cta_step, ++it) {
        const int buf = it & 1;
        mbar_wait(&S.tmem_empty[buf], ((it >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(buf * TC_BN);
        for (int kb = 0; kb < KB; ++kb) {
          mbar_wait(&S.full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a0 = smem_u32(stage_base + stage * STAGE_B);
          const uint32_t b0 = a0 + TC_A_BYTES;
#pragma unroll
          for (int k = 0; k < TC_BK / 16; ++k) {
            const uint32_t off = k * 32;  // 16 bf16 = 32 bytes inside the 128-byte swizzle atom
            if (PAIR) tc_mma_bf16_pair(d_tmem, umma_desc(a0 + off), umma_desc(b0 + off), kIdescBf16Pair, (kb | k) != 0 ? 1u : 0u);
            else tc_mma_bf16(d_tmem, umma_desc(a0 + off), umma_desc(b0 + off), kIdescBf16, (kb | k) != 0 ? 1u : 0u);
          }
          // frees the smem stage once the MMAs above retire, in both CTAs when they share the stage's data
          if (PAIR) tc_commit_pair_mc(&S.empty_bar[stage], (uint16_t)0x3);
          else if (CL == 2) tc_commit_mc(&S.empty_bar[stage], (uint16_t)0x3);
          else tc_commit(&S.empty_bar[stage]);
          if (++stage == NST) { stage = 0; phase ^= 1; }
        }
        if (PAIR) tc_commit_pair_mc(&S.tmem_full[buf], (uint16_t)0x3);   // both CTAs' epilogues drain their half
        else tc_commit(&S.tmem_full[buf]);
      }
    }
  } else {
    // ===================================================================== epilogue warps 2..9
    // Two groups of four warps; group g drains accumulator buffer g, i.e. the tiles with (iteration & 1) == g, so each
    // group has two tile-times per tile.  All per-tile metadata arrives in shared memory through the producer warp's
    // bulk copies: the epilogue issues no global load between the TMEM drain and the survivor append.
    const int q = warp & 3;           // TMEM lane quarter this warp may read
    const int grp = (warp - 2) >> 2;  // 0 or 1
    const bool geo = p.n_constraints > 0;
    const int r = q * 32 + lane;      // accumulator row (candidate) of this thread
    int it = grp;
    for (int t = cta_first + grp * cta_step; t < n_tiles; t += 2 * cta_step, it += 2) {
      const int buf = grp;
      const int ms = it & 3;   // slab set of this tile (same rule as the producer)
      mbar_wait(&S.meta_full[ms], (it >> 2) & 1);
      const VisColMeta* gmeta = S.meta[ms];
      const TcHdr h = S.hdr[ms];
      const VisRowMeta rm = S.rowm[ms][r];
      const int m = h.m0 + r;
      const bool row_ok = m < h.m && rm.ok;
      const int g = h.det_base + m;
      float cx = 0.0f, cy = 0.0f, cr = 0.0f;
      if (geo && row_ok) { cx = f.c_box[(size_t)g * 6]; cy = f.c_box[(size_t)g * 6 + 1]; cr = f.c_radius[g]; }
      // Screen test, E = kScreenRelErr bounds the BF16 operand rounding (|dot~ - dot| <= E |a||b| <= E (|a|^2 + |b|^2) / 2):
      //   cosine: cos >= thr possible   <=>  dot~ >= (thr - 1e-5 - E) |a| * |b|                                = rowk * colb
      //   euclid: d^2 <= thr^2 possible <=>  dot~ >= 0.5 ((1 - 1e-5 - E)(|a|^2 + |b|^2) - thr^2 (1 + 1e-5))   = rowk + colb
      const float rowk = rm.rowk;
      const float* gcolb = S.colb[ms];
      mbar_wait(&S.tmem_full[buf], (it >> 1) & 1);
      tc_fence_after();
      // ---- phase A: drain the accumulator into per-thread survivor masks, then hand the TMEM buffer back at once
      unsigned int keep[TC_BN / 32];
      uint32_t acc[2][32];
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * TC_BN);
      const int nch = min(TC_BN / 32, (h.ncols_left + 31) / 32);   // chunks that hold tracks of this scene
      tc_ld32_issue(taddr, acc[0]);
#pragma unroll
      for (int ch = 0; ch < TC_BN / 32; ++ch) {
        keep[ch] = 0;
        if (ch < nch) {  // warp-uniform
          tc_ld_wait32(acc[ch & 1]);                                     // chunk ch has landed
          if (ch + 1 < nch) tc_ld32_issue(taddr + (ch + 1) * 32, acc[(ch + 1) & 1]);   // in flight while ch is screened
          unsigned int kb = 0;  // bit jj: pair (row, column ch*32+jj) survives the screen
#pragma unroll
          for (int jj = 0; jj < 32; jj += 4) {
            const float4 cb = *reinterpret_cast<const float4*>(gcolb + ch * 32 + jj);
            const float b0 = COSINE ? rowk * cb.x : rowk + cb.x, b1 = COSINE ? rowk * cb.y : rowk + cb.y;
            const float b2 = COSINE ? rowk * cb.z : rowk + cb.z, b3 = COSINE ? rowk * cb.w : rowk + cb.w;
            // a NaN anywhere keeps the pair: the exact pass decides
            if (!(__uint_as_float(acc[ch & 1][jj]) < b0)) kb |= 1u << jj;
            if (!(__uint_as_float(acc[ch & 1][jj + 1]) < b1)) kb |= 2u << jj;
            if (!(__uint_as_float(acc[ch & 1][jj + 2]) < b2)) kb |= 4u << jj;
            if (!(__uint_as_float(acc[ch & 1][jj + 3]) < b3)) kb |= 8u << jj;
          }
          kb &= S.colvalid[ms][ch];   // columns without a usable observation never survive
          const int left = h.ncols_left - ch * 32;   // columns past the scene's last track row hold foreign metadata
          if (left < 32) kb &= (1u << left) - 1u;
          if (!row_ok) kb = 0;
          if (geo && kb) {
            unsigned int kk = kb;
            while (kk) {
              const int jj = __ffs(kk) - 1;
              kk &= kk - 1;
              const VisColGeo cg = colgeo[h.col0 + ch * 32 + jj];   // rare path: straight from global memory
              if (!compat_ok(p, (unsigned int)h.epoch, cg.tep, cx, cy, cr, cg.tx, cg.ty, cg.tr)) kb &= ~(1u << jj);
            }
          }
          keep[ch] = kb;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {   // accumulator buffer drained: the next MMA may start
        if (PAIR) mbar_arrive_cluster(leader_addr(&S.tmem_empty[buf]));
        else mbar_arrive(&S.tmem_empty[buf]);
      }
      // ---- phase B: survivors -> pair list, ONE warp-aggregated append per tile; everything else is None
      int cnt = 0;
#pragma unroll
      for (int ch = 0; ch < TC_BN / 32; ++ch) cnt += __popc(keep[ch]);
      if (__any_sync(0xffffffffu, cnt != 0)) {
        int incl = cnt;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          int tt = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= o) incl += tt;
        }
        const int total = __shfl_sync(0xffffffffu, incl, 31);
        int base = 0;
        if (lane == 31) base = atomicAdd(&f.vis_cnt[h.scene], total);
        base = __shfl_sync(0xffffffffu, base, 31);
        int pos = base + incl - cnt;
#pragma unroll
        for (int ch = 0; ch < TC_BN / 32; ++ch) {
          unsigned int kk = keep[ch];
          while (kk) {
            const int jj = __ffs(kk) - 1;
            kk &= kk - 1;
            if (pos < h.vis_lcap) {
              const VisColMeta cm = gmeta[ch * 32 + jj];
              VisPair vp;
              vp.g = g; vp.row = cm.row; vp.scene = h.scene; vp.outcol = cm.outcol;
              f.vis_pairs[h.vis_lbase + pos] = vp;
            }
            ++pos;
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&S.meta_empty[ms]);   // the slabs of this tile may be overwritten
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CL >= 2) cluster_sync_all();   // no CTA leaves while its peer may still multicast into / arrive on its smem
  if (warp == 1) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// ------------------------------------------------------------------------------------------------ refine kernel
// One warp per surviving pair: f32 distance in the reference's summation order (src/distance.rs:9-47).
__device__ __forceinline__ float reduce_add8_tc(const float* t) {
  float q0 = t[0] + t[4], q1 = t[1] + t[5], q2 = t[2] + t[6], q3 = t[3] + t[7];
  float d0 = q0 + q2, d1 = q1 + q3;
  return d0 + d1;
}

// A warp takes 32 pairs at a time.  Phase 1 (cooperative, coalesced): for every pair the lanes compute the 8-element
// block sums s_blk = reduce_add8(...) of up to 64 blocks and park them in shared memory.  Phase 2: lane p owns pair p and
// adds its block sums strictly in block order, acc = (..((0 + s_0) + s_1) + ..), which is the reference's order; the 32
// serial chains run side by side instead of one 64-step shuffle chain per pair.
constexpr int RF_WARPS = 4;
constexpr int RF_SEG = 64;             // blocks per segment
constexpr int RF_PITCH = RF_SEG + 1;   // +1: lane p reads row p, rows must start in different banks

// a-side of a block sum: the candidate's 8 features of block `blk` (zero padded past D on the TAIL path)
template <bool TAIL>
__device__ __forceinline__ void refine_load_a(const float* __restrict__ a, int blk, int D, float* av) {
  if (!TAIL) {
    const float4 a0 = *reinterpret_cast<const float4*>(a + blk * 8);
    const float4 a1 = *reinterpret_cast<const float4*>(a + blk * 8 + 4);
    av[0] = a0.x; av[1] = a0.y; av[2] = a0.z; av[3] = a0.w; av[4] = a1.x; av[5] = a1.y; av[6] = a1.z; av[7] = a1.w;
  } else {
#pragma unroll
    for (int l = 0; l < 8; ++l) av[l] = blk * 8 + l < D ? a[blk * 8 + l] : 0.0f;   // the input row has D, not d8, floats
  }
}
template <bool COSINE>
__device__ __forceinline__ float refine_block_sum(const float* av, const float* __restrict__ b, int blk) {
  const float4 b0 = *reinterpret_cast<const float4*>(b + blk * 8);
  const float4 b1 = *reinterpret_cast<const float4*>(b + blk * 8 + 4);
  const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
  float t[8];
#pragma unroll
  for (int l = 0; l < 8; ++l) {
    if (COSINE) t[l] = av[l] * bb[l];
    else { const float df = av[l] - bb[l]; t[l] = df * df; }
  }
  return reduce_add8_tc(t);
}

template <bool COSINE, bool TAIL, int RP>
__global__ void __launch_bounds__(RF_WARPS * 32, RP == 32 ? 6 : 8) vis_refine_kernel(Params p, TrackStore ts, Frame f, int* nan_flag) {
  __shared__ float s_bs[RF_WARPS][RP][RF_PITCH];
  const int scene = blockIdx.y;
  if (f.vis_mode[scene] != 0) return;  // survivor list overflowed: this scene is computed densely
  const SceneDesc sc = f.scenes[scene];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int n_pairs = min(f.vis_cnt[scene], sc.vis_lcap);
  const int nblk = p.d8 / 8;
  const int D = p.feature_dim;
  float (*bs)[RF_PITCH] = s_bs[w];
  float vmax = nanf("");
  // warps claim RP survivors at a time from the scene's counter: however many survive, the scene's warps finish together
  for (;;) {
    int i0 = 0;
    if (lane == 0) i0 = atomicAdd(f.refine_next + scene, RP);
    i0 = __shfl_sync(0xffffffffu, i0, 0);
    if (i0 >= n_pairs) break;
    const int npair = min(RP, n_pairs - i0);
    VisPair mine;
    mine.g = 0; mine.row = 0; mine.scene = 0; mine.outcol = 0;
    if (lane < npair) mine = f.vis_pairs[sc.vis_lbase + i0 + lane];
    float acc = 0.0f;
    for (int seg0 = 0; seg0 < nblk; seg0 += RF_SEG) {
      const int segn = min(RF_SEG, nblk - seg0);
      // A candidate's survivors sit next to each other in the list (the observations of its track are adjacent columns
      // of one screen tile; the dense path emits whole groups): its row is loaded once and reused while the candidate
      // stays the same -- a third of the L2 -> SM traffic of this kernel.
      int g_prev = -1;
      float av[RF_SEG / 32][8];
#pragma unroll 4
      for (int pp = 0; pp < npair; ++pp) {
        const int g = __shfl_sync(0xffffffffu, mine.g, pp);
        const int row = __shfl_sync(0xffffffffu, mine.row, pp);
        const float* b = ts.feat + (size_t)row * p.d8;
        if (g != g_prev) {   // warp-uniform
          const float* a = f.in_feat + (size_t)g * D;
#pragma unroll
          for (int h = 0; h < RF_SEG / 32; ++h)
            if (h * 32 + lane < segn) refine_load_a<TAIL>(a, seg0 + h * 32 + lane, D, av[h]);
          g_prev = g;
        }
#pragma unroll
        for (int h = 0; h < RF_SEG / 32; ++h) {
          const int j = h * 32 + lane;
          if (j < segn) bs[pp][j] = refine_block_sum<COSINE>(av[h], b, seg0 + j);
        }
      }
      __syncwarp();
      if (lane < npair)
        for (int j = 0; j < segn; ++j) acc = acc + bs[lane][j];
      __syncwarp();
    }
    if (lane < npair) {
      float v = nanf("");
      if (COSINE) {
        const float d = acc / sqrtf(f.c_norm2[mine.g] * ts.fnorm2[mine.row]);
        if (d >= p.visual_threshold) v = 1.0f - d;       // is_ok + distance_to_weight
      } else {
        const float d = sqrtf(acc);
        if (d <= p.visual_threshold) v = d;
      }
      f.vis_val[sc.vis_lbase + i0 + lane] = v;
      if (!is_nan(v) && !(v <= vmax)) vmax = v;   // best.rs "max_dist": maximum over the entries that exist
      if (nan_flag && is_nan(v)) nan_flag[scene] = 1;   // dense path: an entry the threshold cuts voids its precondition
    }
  }
  // one atomic per warp
  unsigned int u = 0u;
  if (!is_nan(vmax)) {
    u = __float_as_uint(vmax);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) u = max(u, __shfl_xor_sync(0xffffffffu, u, o));
  if (lane == 0 && u != 0u) atomicMax(f.scene_max + scene, u);
}

// ------------------------------------------------------------------------------------------------ refine, asynchronous copies
// Same arithmetic, different data movement.  The register-staged kernel above keeps about 4 KB per warp in flight and has
// none in flight while a warp adds its serial chains; ncu (round 1) showed it latency-bound (DRAM 47 %, 23 % of the warp
// slots active).  Here every warp owns a ring of three shared-memory stages and feeds it with cp.async (16-byte chunks, no
// registers held): while it turns the rows of stage s into block sums, the copies of stages s+1 and s+2 -- 16 KB -- are in
// flight, whatever the warp is doing.  A stage holds the 512-float segments of both rows of two pairs; chunks are stored
// half-block-major so that the 16-byte reads of the block sums hit 32 different banks.  The 64 block sums of 32 pairs are
// parked as before and added in the reference's order by 32 lanes side by side.
constexpr int RA_WARPS = 6;
constexpr int RA_STAGES = 3;
constexpr int RA_PAIRS = 2;                        // pairs per stage
constexpr int RA_SEGF = RF_SEG * 8;                // floats per row segment (512)
constexpr int RA_STAGE_FLOATS = RA_PAIRS * 2 * RA_SEGF;   // 2048 floats = 8 KB

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <bool COSINE>
__global__ void __launch_bounds__(RA_WARPS * 32) vis_refine_async_kernel(Params p, TrackStore ts, Frame f, int* nan_flag) {
  extern __shared__ __align__(16) unsigned char ra_smem[];
  const int scene = blockIdx.y;
  if (f.vis_mode[scene] != 0) return;  // survivor list overflowed: this scene is computed densely
  const SceneDesc sc = f.scenes[scene];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  float* stage0 = reinterpret_cast<float*>(ra_smem) + (size_t)w * (RA_STAGES * RA_STAGE_FLOATS + 32 * RF_PITCH);
  float (*bs)[RF_PITCH] = reinterpret_cast<float (*)[RF_PITCH]>(stage0 + RA_STAGES * RA_STAGE_FLOATS);
  const int n_pairs = min(f.vis_cnt[scene], sc.vis_lcap);
  const int nblk = p.d8 / 8;
  const int D = p.feature_dim;   // == d8 on this path, rows 16-byte aligned
  float vmax = nanf("");
  for (;;) {
    int i0 = 0;
    if (lane == 0) i0 = atomicAdd(f.refine_next + scene, 32);
    i0 = __shfl_sync(0xffffffffu, i0, 0);
    if (i0 >= n_pairs) break;
    const int npair = min(32, n_pairs - i0);
    VisPair mine;
    mine.g = 0; mine.row = 0; mine.scene = 0; mine.outcol = 0;
    if (lane < npair) mine = f.vis_pairs[sc.vis_lbase + i0 + lane];
    float acc = 0.0f;
    const int nst = (npair + RA_PAIRS - 1) / RA_PAIRS;   // stages of this batch
    for (int seg0 = 0; seg0 < nblk; seg0 += RF_SEG) {
      const int segn = min(RF_SEG, nblk - seg0);          // blocks of this segment
      const int nchunk = segn * 2;                        // 16-byte chunks per row segment
      auto issue = [&](int st) {
        if (st < nst) {
          float* dst = stage0 + (st % RA_STAGES) * RA_STAGE_FLOATS;
#pragma unroll
          for (int pi = 0; pi < RA_PAIRS; ++pi) {
            const int pp = st * RA_PAIRS + pi;
            const int g = __shfl_sync(0xffffffffu, mine.g, pp & 31);
            const int row = __shfl_sync(0xffffffffu, mine.row, pp & 31);
            if (pp < npair) {
              const float* a = f.in_feat + (size_t)g * D + seg0 * 8;
              const float* b = ts.feat + (size_t)row * p.d8 + seg0 * 8;
              float* da = dst + (pi * 2) * RA_SEGF;
              float* db = da + RA_SEGF;
#pragma unroll
              for (int c = 0; c < RA_SEGF / 4 / 32; ++c) {
                const int q = c * 32 + lane;   // chunk q = half (q & 1) of block (q >> 1); stored half-major
                if (q < nchunk) {
                  const int pos = ((q & 1) * RF_SEG + (q >> 1)) * 4;
                  cp_async16(da + pos, a + q * 4);
                  cp_async16(db + pos, b + q * 4);
                }
              }
            }
          }
        }
        cp_async_commit();   // always: the group count stays in step with the stage count
      };
      issue(0);
      issue(1);
      for (int st = 0; st < nst; ++st) {
        issue(st + 2);
        cp_async_wait<2>();
        __syncwarp();
        const float* src = stage0 + (st % RA_STAGES) * RA_STAGE_FLOATS;
#pragma unroll
        for (int pi = 0; pi < RA_PAIRS; ++pi) {
          const int pp = st * RA_PAIRS + pi;
          if (pp < npair) {   // warp-uniform
            const float* sa = src + (pi * 2) * RA_SEGF;
            const float* sb2 = sa + RA_SEGF;
#pragma unroll
            for (int h = 0; h < RF_SEG / 32; ++h) {
              const int j = h * 32 + lane;
              if (j < segn) {
                const float4 a0 = *reinterpret_cast<const float4*>(sa + j * 4);
                const float4 a1 = *reinterpret_cast<const float4*>(sa + (RF_SEG + j) * 4);
                const float4 b0 = *reinterpret_cast<const float4*>(sb2 + j * 4);
                const float4 b1 = *reinterpret_cast<const float4*>(sb2 + (RF_SEG + j) * 4);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float t[8];
#pragma unroll
                for (int l = 0; l < 8; ++l) {
                  if (COSINE) t[l] = av[l] * bv[l];
                  else { const float df = av[l] - bv[l]; t[l] = df * df; }
                }
                bs[pp][j] = reduce_add8_tc(t);
              }
            }
          }
        }
        __syncwarp();   // the stage may be refilled (two iterations ahead), the block sums are visible
      }
      cp_async_wait<0>();
      if (lane < npair)
        for (int j = 0; j < segn; ++j) acc = acc + bs[lane][j];
      __syncwarp();
    }
    if (lane < npair) {
      float v = nanf("");
      if (COSINE) {
        const float d = acc / sqrtf(f.c_norm2[mine.g] * ts.fnorm2[mine.row]);
        if (d >= p.visual_threshold) v = 1.0f - d;       // is_ok + distance_to_weight
      } else {
        const float d = sqrtf(acc);
        if (d <= p.visual_threshold) v = d;
      }
      f.vis_val[sc.vis_lbase + i0 + lane] = v;
      if (!is_nan(v) && !(v <= vmax)) vmax = v;   // best.rs "max_dist": maximum over the entries that exist
      if (nan_flag && is_nan(v)) nan_flag[scene] = 1;
    }
  }
  unsigned int u = 0u;
  if (!is_nan(vmax)) {
    u = __float_as_uint(vmax);
    u = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) u = max(u, __shfl_xor_sync(0xffffffffu, u, o));
  if (lane == 0 && u != 0u) atomicMax(f.scene_max + scene, u);
}

// dense view of the sparse scenes' visual entries (operators / debugging): None everywhere, then the refined survivors
__global__ void vis_fill_none_kernel(Params p, Frame f) {
  const int scene = blockIdx.y;
  if (f.scene_mode[scene] != 0) return;
  const SceneDesc sc = f.scenes[scene];
  const long long cnt = (long long)sc.m * sc.n * p.max_obs;
  float* out = f.vis + sc.vis_off;
  const float qnan = nanf("");
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < cnt; i += (long long)gridDim.x * blockDim.x) out[i] = qnan;
}
__global__ void vis_scatter_kernel(Params p, Frame f) {
  const int scene = blockIdx.y;
  if (f.scene_mode[scene] != 0) return;
  const SceneDesc sc = f.scenes[scene];
  const int n_pairs = min(f.vis_cnt[scene], sc.vis_lcap);
  float* out = f.vis + sc.vis_off;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n_pairs; i += gridDim.x * blockDim.x) {
    const VisPair vp = f.vis_pairs[sc.vis_lbase + i];
    out[(size_t)(vp.g - sc.det_base) * (sc.n * p.max_obs) + vp.outcol] = f.vis_val[sc.vis_lbase + i];
  }
}
void launch_vis_densify(const Params& p, const Frame& f, int n_scenes, cudaStream_t st) {
  if (n_scenes == 0) return;
  dim3 grid(64, n_scenes);
  vis_fill_none_kernel<<<grid, 256, 0, st>>>(p, f);
  vis_scatter_kernel<<<grid, 256, 0, st>>>(p, f);
  note_launch(2);
}

// ------------------------------------------------------------------------------------------------ bf16 operand copies
__global__ void to_bf16_kernel(const float* src, int src_pitch, int d, int d8, long long rows, __nv_bfloat16* dst) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * d8) return;
  long long r = i / d8;
  int c = (int)(i - r * d8);
  float x = c < d ? src[r * src_pitch + c] : 0.0f;
  dst[i] = __float2bfloat16_rn(x);
}

void launch_to_bf16(const float* src, int src_pitch, int d, int d8, long long rows, void* dst, cudaStream_t st) {
  if (rows == 0) return;
  long long n = rows * d8;
  to_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, src_pitch, d, d8, rows, (__nv_bfloat16*)dst);
  note_launch();
}

// ------------------------------------------------------------------------------------------------ host launcher
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qr) == cudaSuccess && p) fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_map(CUtensorMap* m, const void* base, long long rows, int d8, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return -1;
  cuuint64_t dims[2] = {(cuuint64_t)d8, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)d8 * 2};
  cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -2;
}

// per-frame metadata: one thread per physical feature row (scene, arena block b, physical slot p) and per candidate.
// A block without an owner (free list) is an invalid column; otherwise the row belongs to track n = blk_owner[b].
__global__ void vis_meta_kernel(Params p, TrackStore ts, Frame f, int n_scenes, int max_rows, VisColMeta* colmeta,
                                VisColGeo* colgeo, float* colb, unsigned int* colvalid) {
  const int s = blockIdx.y;
  const SceneDesc sc = f.scenes[s];
  const int K = p.max_obs;
  const int prow = blockIdx.x * blockDim.x + threadIdx.x;
  const bool in = prow < sc.nb * K && prow < max_rows;
  VisColMeta cm;
  cm.colb = 0.0f; cm.colc = 0.0f; cm.outcol = -1; cm.row = -1;
  if (in) {
    const int b = prow / K, ph = prow - b * K;
    const size_t sbase = (size_t)sc.slot * ts.track_cap;
    const int n = ts.blk_owner ? ts.blk_owner[sbase + b] : b;
    const size_t ti = sbase + (n >= 0 ? n : 0);
    const size_t frow = (sbase + b) * K + ph;   // feature row of this column
    unsigned int tep = 0u;
    if (n >= 0) {
      const int on = ts.obs_n[ti];
      // logical <-> physical observation bookkeeping of this track
      int k_of = -1, live_mask = 0;
      for (int k = 0; k < K; ++k) {
        if (k < on && ts.obs_hasf[ti * K + k]) {
          int pp = ts.obs_phys[ti * K + k];
          live_mask |= 1 << pp;
          if (pp == ph) k_of = k;
        }
      }
      tep = ts.epoch[ti];
      if (k_of >= 0) {
        const unsigned int delta = sc.epoch > tep ? sc.epoch - tep : tep - sc.epoch;
        cm.outcol = n * K + k_of;
        const bool valid = (ts.feat_cnt[ti] >= p.min_track_length) && ((unsigned int)p.max_idle_epochs >= delta);
        const float nb = ts.fnorm2[frow];
        cm.colb = p.visual_kind == 1 ? sqrtf(nb) : 0.5f * nb * (1.0f - 1e-5f - kScreenRelErr);
        cm.row = valid ? (int)frow : -1;
      } else {
        // dead physical slot -> owns the dead_rank-th logical column without a feature (written as None)
        int dead_rank = 0;
        for (int pp = 0; pp < ph; ++pp) dead_rank += ((live_mask >> pp) & 1) ? 0 : 1;
        int seen = 0;
        for (int k = 0; k < K; ++k) {
          bool lv = k < on && ts.obs_hasf[ti * K + k];
          if (!lv) { if (seen == dead_rank) { cm.outcol = n * K + k; break; } ++seen; }
        }
      }
    }
    colmeta[sc.col_off + prow] = cm;
    colb[sc.col_off + prow] = cm.colb;
    if (p.n_constraints > 0) {
      VisColGeo cg;
      cg.tx = 0.0f; cg.ty = 0.0f; cg.tr = 0.0f; cg.tep = tep;
      if (n >= 0) { const float* tb = ts.pred + ti * 6; cg.tx = tb[0]; cg.ty = tb[1]; cg.tr = ts.radius[ti]; }
      colgeo[sc.col_off + prow] = cg;
    }
  }
  // col_off is a multiple of 128, so the 32 columns of a warp are exactly one word of the validity mask
  const unsigned int vb = __ballot_sync(0xffffffffu, cm.row >= 0);
  if ((threadIdx.x & 31) == 0 && vb != 0) colvalid[(sc.col_off + prow) >> 5] = vb;
  else if ((threadIdx.x & 31) == 0 && in) colvalid[(sc.col_off + prow) >> 5] = 0u;
}

__global__ void vis_rowmeta_kernel(Params p, Frame f, VisRowMeta* rowmeta) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= f.total) return;
  g += f.det0;
  const float na = f.c_norm2[g];
  VisRowMeta rm;
  rm.ok = (f.c_flags[g] & 2) ? 1 : 0;
  const float thr = p.visual_threshold;
  if (p.visual_kind == 1) rm.rowk = (thr - 1e-5f - kScreenRelErr) * sqrtf(na);
  else rm.rowk = 0.5f * (na * (1.0f - 1e-5f - kScreenRelErr) - thr * thr * (1.0f + 1e-5f));
  rowmeta[g] = rm;
}

// exact refinement of the scene pair lists (f.vis_pairs / vis_cnt / refine_next / vis_val; scenes with vis_mode != 0 skipped)
int launch_vis_refine(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int* nan_flag, cudaStream_t st) {
  if (n_scenes == 0) return 0;
  dim3 grid(16, n_scenes);   // 64 warps x 32 survivors per scene in flight; more survivors are claimed in further rounds
  // the vector path needs 16-byte aligned input rows (a caller-owned device pointer on the device-io path)
  const bool tail = p.feature_dim != p.d8 || (reinterpret_cast<uintptr_t>(f.in_feat) & 15) != 0;
  // SB200_REFINE=async: the cp.async variant (measured slower on B200: 0.39 vs 0.25 ms at cfg5 -- one CTA of six warps per
  // SM cannot keep the block-sum arithmetic fed; kept for experiments)
  static const bool async_copy = getenv("SB200_REFINE") != nullptr && !strcmp(getenv("SB200_REFINE"), "async");
  if (!tail && async_copy) {
    // asynchronous-copy kernel: 3 x 8 KB stages + the parked block sums per warp
    const size_t smem = (size_t)RA_WARPS * (RA_STAGES * RA_STAGE_FLOATS + 32 * RF_PITCH) * sizeof(float);
    const void* fn = p.visual_kind == 1 ? (const void*)vis_refine_async_kernel<true> : (const void*)vis_refine_async_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return (int)e;
    dim3 grid2(8, n_scenes);
    if (p.visual_kind == 1) vis_refine_async_kernel<true><<<grid2, RA_WARPS * 32, smem, st>>>(p, ts, f, nan_flag);
    else vis_refine_async_kernel<false><<<grid2, RA_WARPS * 32, smem, st>>>(p, ts, f, nan_flag);
    note_launch();
    return 0;
  }
  // 16 survivors per claim: half the parked block sums of 32, eight CTAs per SM instead of six (0.182 vs 0.200 ms at cfg5);
  // SB200_REFINE_PAIRS=32 selects the wider claim
  static const bool rp16 = !(getenv("SB200_REFINE_PAIRS") != nullptr && atoi(getenv("SB200_REFINE_PAIRS")) == 32);
#define SB_RF(C, T)                                                                        \
  do {                                                                                     \
    if (rp16) vis_refine_kernel<C, T, 16><<<grid, RF_WARPS * 32, 0, st>>>(p, ts, f, nan_flag); \
    else vis_refine_kernel<C, T, 32><<<grid, RF_WARPS * 32, 0, st>>>(p, ts, f, nan_flag);      \
  } while (0)
  if (p.visual_kind == 1) {
    if (tail) SB_RF(true, true); else SB_RF(true, false);
  } else {
    if (tail) SB_RF(false, true); else SB_RF(false, false);
  }
#undef SB_RF
  note_launch();
  return 0;
}

void launch_vis_colmeta(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_n, const TcArgs& tc,
                        cudaStream_t st) {
  const int max_rows = tc.max_rows > 0 ? tc.max_rows : max_n * p.max_obs;
  if (tc.n_tiles == 0 || max_rows <= 0) return;
  dim3 grid((max_rows + 255) / 256, n_scenes);
  vis_meta_kernel<<<grid, 256, 0, st>>>(p, ts, f, n_scenes, max_rows, tc.colmeta, tc.colgeo, tc.colb, tc.colvalid);
  note_launch();
}

int launch_vis_cost_tc(const Params& p, const TrackStore& ts, const Frame& f, int n_scenes, int max_n, const TcArgs& tc,
                       int phase, cudaStream_t st) {
  if (tc.n_tiles == 0) return 0;
  if (phase == 1) {
    int rc = launch_vis_refine(p, ts, f, n_scenes, nullptr, st);
    if (tc.ev_refine1) cudaEventRecord(tc.ev_refine1, st);
    return rc;
  }
  const bool cluster = tc.cluster2;
  CUtensorMap mA, mB;
  if (make_map(&mA, f.c_bf16, tc.a_rows, p.d8, TC_BM) || make_map(&mB, ts.feat_bf16, tc.b_rows, p.d8, cluster ? TC_BN / 2 : TC_BN))
    return -1;
  size_t smem = sizeof(TcSmem) + 1024;
  const bool cosine = p.visual_kind == 1;
  cudaError_t e = cudaSuccess;
  const void* fn = tc.pair ? (cosine ? (const void*)vis_screen_kernel<3, true> : (const void*)vis_screen_kernel<3, false>)
                   : cluster ? (cosine ? (const void*)vis_screen_kernel<2, true> : (const void*)vis_screen_kernel<2, false>)
                             : (cosine ? (const void*)vis_screen_kernel<1, true> : (const void*)vis_screen_kernel<1, false>);
  e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  if (!tc.colmeta_done) launch_vis_colmeta(p, ts, f, n_scenes, max_n, tc, st);
  vis_rowmeta_kernel<<<(f.total + 255) / 256, 256, 0, st>>>(p, f, tc.rowmeta);
  note_launch();
  if (tc.ev_screen0) cudaEventRecord(tc.ev_screen0, st);
  {
    const int ncta = cluster ? 2 * std::min(tc.n_tiles, tc.num_sms / 2) : std::min(tc.n_tiles, tc.num_sms);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(ncta);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cluster ? 2 : 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    const TcTile* d_tiles = tc.d_tiles;
    int n_tiles = tc.n_tiles;
    const int* d_n_tiles = tc.d_n_tiles;
    const VisColMeta* cmeta = tc.colmeta;
    const VisColGeo* cgeo = tc.colgeo;
    const VisRowMeta* rmeta = tc.rowmeta;
    const float* cb = tc.colb;
    const unsigned int* cv = tc.colvalid;
    void* args[] = {(void*)&mA, (void*)&mB, (void*)&p, (void*)&ts, (void*)&f, (void*)&d_tiles, (void*)&n_tiles,
                    (void*)&d_n_tiles, (void*)&cmeta, (void*)&cgeo, (void*)&rmeta, (void*)&cb, (void*)&cv};
    e = cudaLaunchKernelExC(&cfg, fn, args);
    if (e != cudaSuccess) return (int)e;
    note_launch();
  }
  if (tc.ev_screen1) cudaEventRecord(tc.ev_screen1, st);
  return 0;
}

}  // namespace sb
