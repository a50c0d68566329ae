// tcgen05 (5th-gen tensor core) GEMM with error-compensated 3xTF32 arithmetic -- the fast path of the linear
// layers of gcbf.nn.MLP (reference gcbf/nn/mlp.py:44-47; 2048-wide phi / gamma GEMMs are >99 % of the FLOPs).
//
//   D[Mo, No] = sum_k A[Mo, k] * B[No, k]          (both operands K-major fp32)
//
// fp32-grade accuracy on the tensor cores (plain TF32 misses the 1e-5 parity bar): every operand x
// is split into hi = tf32(x) and lo = x - hi (exact), and the product is accumulated in fp32 TMEM as
//        hi*hi + lo*hi + hi*lo          (3 tcgen05.mma.kind::tf32 per k-slice; lo*lo ~ 2^-22 is dropped)
//
// Round-1 structure (one code path for forward / data-grad / weight-grad):
//   1. `split_kernel` writes the (optionally transposed) operand as [hi; lo] K-major scratch, zero padded to
//      multiples of the tile, so the three reference layouts all become the same TN problem and the TMA never
//      sees a ragged edge;
//   2. `gemm_tc_kernel`: persistent CTAs, warp-specialised -- warp 0 = TMA producer (cp.async.bulk.tensor,
//      128B-swizzled 128x32 / BNx32 fp32 boxes, 4 per stage), warp 1 = MMA issuer (one thread issues
//      3 x BK/8 tcgen05.mma per stage, accumulators double-buffered in TMEM), warps 2-5 = epilogue
//      (tcgen05.ld 32x32b.x32 -> fused alpha/bias/activation | ReLU-mask | accumulate -> global).
#include <cuda.h>
#include <stdlib.h>

#include "common.cuh"

namespace gcbf {

namespace tc {

constexpr int BM = 128;
#ifndef GCBF_TC_BK
#define GCBF_TC_BK 16
#endif
constexpr int BK = GCBF_TC_BK;         // fp32 elements per k-block: 32 = 128 B rows (SWIZZLE_128B), 16 = 64 B rows (SWIZZLE_64B)
constexpr int SWIZZLE_BYTES = BK * 4;  // one smem row of a tile == one swizzle span
static_assert(BK == 16 || BK == 32, "BK must be 16 or 32");
constexpr int UMMA_K = 8;              // tf32: 32 bytes per instruction
constexpr int NUM_THREADS = 320;       // warp0 TMA, warp1 MMA, warps 2..9 epilogue / fp32 chunk accumulation
constexpr int KCH = 256 / BK;          // k-blocks accumulated inside the tensor core (256 K-elements) before promotion to registers
#ifndef GCBF_TC_CONV_GROUPS
#define GCBF_TC_CONV_GROUPS 4
#endif
constexpr int CONV_GROUPS = GCBF_TC_CONV_GROUPS;   // epilogue warps form this many converter groups taking k-blocks round robin
constexpr int MAX_CHUNK_ROWS = 65536;  // rows of the big operand processed per launch (bounds the scratch)

enum { EPI_FWD = 0, EPI_DGRAD = 1, EPI_WGRAD = 2 };

struct EpiParams {
  int mode;
  const float* alpha;
  const float* bias;
  int act;
  const float* relu_src;
  int ld_relu;
  int accumulate;
  int atomic;
  int dbg;   // experiment switches (GCBF_TC_DBG): 1 = converters skip the split, 2 = write lo only
};

// ---- PTX wrappers ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ uint32_t mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}

__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void tmem_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// D[tmem] (+)= A[smem desc] * B[smem desc], tf32 inputs, fp32 accumulate
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// arrive on an mbarrier once all previously issued MMAs of this thread have completed
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// UMMA shared-memory descriptor, K-major, SWIZZLE_128B (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start address >> 4 | [16,30) leading byte offset >> 4 (unused for swizzled K-major: 1) |
//   [32,46) stride byte offset >> 4 (8 rows x 128 B = 1024 B) | [46,48) version = 1 | [61,64) layout = 2 (SW128)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((8 * SWIZZLE_BYTES) >> 4) << 32;   // stride between 8-row swizzle atoms
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)(SWIZZLE_BYTES == 128 ? 2 : 4) << 61;  // SWIZZLE_128B = 2, SWIZZLE_64B = 4
  return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 (1) @4, a/b format TF32 (2) @7/@10,
// a/b major K (0) @15/@16, N>>3 @17, M>>4 @24
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---- operand preparation: [hi; lo] split, optional transpose, zero padding -----------------------------------
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  const uint32_t b = __float_as_uint(x);
  if ((b & 0x7f800000u) == 0x7f800000u) { hi = x; lo = 0.f; return; }   // inf / nan pass through
  hi = __uint_as_float((b + 0x1000u) & 0xffffe000u);                     // round to nearest tf32 (10-bit mantissa)
  lo = __fsub_rn(x, hi);                                                 // exact
}

// dst_hi[r][c] (r < R_pad, c < C_pad), element (r, c) = TRANS ? src[c][r] : src[r][c]; zero outside [rows, cols)
template <int BN>
struct Cfg {
  static constexpr int STAGES = ((BN == 256) ? 2 : 3) * (32 / BK);
  static constexpr int A_BYTES = BM * BK * 4;          // one of hi / lo
  static constexpr int B_BYTES = BN * BK * 4;
  static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
  static constexpr int TMEM_COLS = 2 * BN;             // double-buffered accumulator (power of two: 256 / 512)
};

// A_RAW = true: operand A is the caller's fp32 matrix itself (TMA straight from it, out-of-range rows / columns zero
// filled); the epilogue warps split every landed 128x32 tile into hi (in place) and lo in shared memory before the MMA
// warp may read it -- no [hi; lo] scratch copy of the big activation operand in HBM.
template <int BN, bool A_RAW>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b, int a_lo_row,
               int b_lo_row, float* __restrict__ C, int ldc, int Mo, int No, int tiles_m, int tiles_n,
               int kblocks_per_split, int kblocks_total, EpiParams ep) {
  using K = Cfg<BN>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::STAGES * K::STAGE_BYTES);
  uint64_t* full = bars;                      // [STAGES]  TMA -> MMA
  uint64_t* empty = bars + K::STAGES;         // [STAGES]  MMA -> TMA
  uint64_t* tfull = bars + 2 * K::STAGES;     // [2]       MMA -> epilogue
  uint64_t* tempty = bars + 2 * K::STAGES + 2;  // [2]     epilogue -> MMA
  uint64_t* conv = bars + 2 * K::STAGES + 4;    // [STAGES] converters -> MMA (A_RAW only)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3 * K::STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kb0 = blockIdx.y * kblocks_per_split;
  const int kb1 = min(kblocks_total, kb0 + kblocks_per_split);
  const int nkb = kb1 - kb0;
  const int num_tiles = tiles_m * tiles_n;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&map_a);
    tma_prefetch_desc(&map_b);
    for (int s = 0; s < K::STAGES; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); mbar_init(&conv[s], 8 / CONV_GROUPS); }
    for (int b = 0; b < 2; ++b) { mbar_init(&tfull[b], 1); mbar_init(&tempty[b], 8 * 32); }
    fence_barrier_init();
  }
  if (warp == 1) {            // one warp allocates TMEM and later frees it
    tmem_alloc(tmem_slot, K::TMEM_COLS);
    tmem_relinquish();
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (nkb > 0) {
    if (warp == 0) {
      // ===== TMA producer =====
      if (lane == 0) {
        int stage = 0;
        uint32_t phase = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
          const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
          for (int kb = kb0; kb < kb1; ++kb) {
            mbar_wait(&empty[stage], phase ^ 1);
            uint8_t* st = smem + stage * K::STAGE_BYTES;
            mbar_expect_tx(&full[stage], A_RAW ? K::STAGE_BYTES - K::A_BYTES : K::STAGE_BYTES);
            tma_load_2d(st, &map_a, &full[stage], kb * BK, m0);
            if (!A_RAW) tma_load_2d(st + K::A_BYTES, &map_a, &full[stage], kb * BK, a_lo_row + m0);
            tma_load_2d(st + 2 * K::A_BYTES, &map_b, &full[stage], kb * BK, n0);
            tma_load_2d(st + 2 * K::A_BYTES + K::B_BYTES, &map_b, &full[stage], kb * BK, b_lo_row + n0);
            if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
          }
        }
      }
    } else if (warp == 1) {
      // ===== MMA issuer (single thread) =====
      // The tensor core's fp32 accumulator truncates on every MMA (tools/acc_probe.py: bias ~ -K/8 * 2^-24), so K is
      // consumed in chunks of KCH*32 = 256 elements: each chunk starts a fresh TMEM accumulator (double-buffered) and the
      // epilogue warps add the chunk sums in registers with round-to-nearest.
      if (lane == 0) {
        constexpr uint32_t idesc = make_idesc(BM, BN);
        int stage = 0;
        uint32_t phase = 0;
        int buf = 0;
        uint32_t tphase[2] = {0, 0};
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
          for (int kc = 0; kc < nkb; kc += KCH) {
            mbar_wait(&tempty[buf], tphase[buf] ^ 1);        // epilogue has drained this accumulator
            tcgen05_fence_after();
            const uint32_t d_tmem = tmem_base + (uint32_t)(buf * BN);
            const int kend = min(nkb, kc + KCH);
            for (int kb = kc; kb < kend; ++kb) {
              mbar_wait(&full[stage], phase);                 // TMA bytes (A and B) have landed
              if (A_RAW) mbar_wait(&conv[stage], phase);      // ... and A has been split into hi / lo
              tcgen05_fence_after();
              const uint32_t st = smem_u32(smem + stage * K::STAGE_BYTES);
              const uint64_t a_hi = make_smem_desc(st), a_lo = make_smem_desc(st + K::A_BYTES);
              const uint64_t b_hi = make_smem_desc(st + 2 * K::A_BYTES), b_lo = make_smem_desc(st + 2 * K::A_BYTES + K::B_BYTES);
#pragma unroll
              for (int kk = 0; kk < BK / UMMA_K; ++kk) {
                const uint64_t adv = (uint64_t)((kk * UMMA_K * 4) >> 4);   // +32 B per k-slice inside the 128 B swizzle row
                umma_tf32(d_tmem, a_lo + adv, b_hi + adv, idesc, (kb > kc || kk > 0) ? 1u : 0u);
                umma_tf32(d_tmem, a_hi + adv, b_lo + adv, idesc, 1u);
                umma_tf32(d_tmem, a_hi + adv, b_hi + adv, idesc, 1u);
              }
              umma_commit(&empty[stage]);                     // smem slot free once these MMAs retire
              if (++stage == K::STAGES) { stage = 0; phase ^= 1; }
            }
            umma_commit(&tfull[buf]);                          // chunk sum complete -> epilogue
            tphase[buf] ^= 1;
            buf ^= 1;
          }
        }
      }
    } else {
      // ===== epilogue warps 2..9: TMEM lane quarter = warp % 4, column half = (warp - 2) / 4 =====
      constexpr int CH = BN / 2;                               // columns owned by one thread
      const int lg = warp & 3;
      const int chalf = (warp - 2) >> 2;
      const float alpha = ep.alpha ? __ldg(ep.alpha) : 1.f;
      int buf = 0;
      uint32_t tphase[2] = {0, 0};
      uint32_t gk_base = 0;   // k-blocks this CTA has consumed so far (all tiles): stage = gk % STAGES
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int m0 = (t / tiles_n) * BM, n0 = (t % tiles_n) * BN;
        float acc[CH];
#pragma unroll
        for (int j = 0; j < CH; ++j) acc[j] = 0.f;
        const int nch = (nkb + KCH - 1) / KCH;
        int promoted = 0;
        // add the finished chunk sum of TMEM buffer `buf` into the register accumulators (round-to-nearest fp32)
        auto promote = [&]() {
          mbar_wait(&tfull[buf], tphase[buf]);
          tcgen05_fence_after();
          const uint32_t taddr = tmem_base + (uint32_t)(buf * BN + chalf * CH) + ((uint32_t)(lg * 32) << 16);
#pragma unroll
          for (int c = 0; c < CH / 32; ++c) {
            uint32_t r[32];
            tmem_ld_32x32b_x32(taddr + (uint32_t)(c * 32), r);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[c * 32 + j] = __fadd_rn(acc[c * 32 + j], __uint_as_float(r[j]));
          }
          tcgen05_fence_before();
          mbar_arrive(&tempty[buf]);
          tphase[buf] ^= 1;
          buf ^= 1;
          ++promoted;
        };
        if (A_RAW) {
          // converter duty: split each landed A tile into hi (in place) and lo.  The 8 warps work as CONV_GROUPS groups
          // that take k-blocks round robin (each group has CONV_GROUPS MMA k-block periods per tile it converts); addressing is
          // explicit shared-space PTX (16-byte chunks, consecutive threads -> consecutive chunks: conflict free).
          constexpr int GT = 256 / CONV_GROUPS;                        // threads per converter group
          const int grp = (warp - 2) / (GT / 32);
          const int gt = ((warp - 2) % (GT / 32)) * 32 + lane;         // 0..GT-1 inside the group
          constexpr int CHUNKS = K::A_BYTES / 16;                      // 16-byte chunks per A tile
          static_assert(CHUNKS % GT == 0, "converter mapping");
          for (int kb = 0; kb < nkb; ++kb) {
            const uint32_t gk = gk_base + (uint32_t)kb;
            if ((int)(gk % (uint32_t)CONV_GROUPS) == grp) {
              const int cs = (int)(gk % (uint32_t)K::STAGES);
              mbar_wait(&full[cs], (gk / (uint32_t)K::STAGES) & 1u);
              const uint32_t a_hi = smem_u32(smem + cs * K::STAGE_BYTES);
              if (!(ep.dbg & 1)) {
#pragma unroll
                for (int q = 0; q < CHUNKS / GT; ++q) {
                  const uint32_t addr = a_hi + (uint32_t)((q * GT + gt) * 16);
                  uint32_t x0, x1, x2, x3;
                  asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(x0), "=r"(x1), "=r"(x2), "=r"(x3) : "r"(addr));
                  const uint32_t rnd = (ep.dbg & 4) ? 0u : 0x1000u;   // dbg 4: truncate (what the tensor core would do to raw x)
                  const uint32_t h0 = (x0 + rnd) & 0xffffe000u, h1 = (x1 + rnd) & 0xffffe000u;
                  const uint32_t h2 = (x2 + rnd) & 0xffffe000u, h3 = (x3 + rnd) & 0xffffe000u;
                  const uint32_t l0 = __float_as_uint(__fsub_rn(__uint_as_float(x0), __uint_as_float(h0)));
                  const uint32_t l1 = __float_as_uint(__fsub_rn(__uint_as_float(x1), __uint_as_float(h1)));
                  const uint32_t l2 = __float_as_uint(__fsub_rn(__uint_as_float(x2), __uint_as_float(h2)));
                  const uint32_t l3 = __float_as_uint(__fsub_rn(__uint_as_float(x3), __uint_as_float(h3)));
                  if (!(ep.dbg & 4))
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(h0), "r"(h1), "r"(h2), "r"(h3) : "memory");
                  asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(addr + (uint32_t)K::A_BYTES), "r"(l0), "r"(l1), "r"(l2), "r"(l3)
                               : "memory");
                }
              }
              fence_proxy_async();          // generic-proxy smem writes -> visible to the tensor core (async proxy)
              __syncwarp();
              if (lane == 0) mbar_arrive(&conv[cs]);
            }
            // chunk c is complete once the MMA warp is past k-block KCH*(c+1) - 1; by the time this warp sees k-block
            // KCH*(c+1) + STAGES the MMA warp has released that block's stage, so the wait below does not stall
            if (kb >= KCH + K::STAGES && (kb - K::STAGES) % KCH == 0) promote();
          }
          gk_base += (uint32_t)nkb;
          while (promoted < nch) promote();
        } else {
          for (int c = 0; c < nch; ++c) promote();
        }
        const int row = m0 + lg * 32 + lane;
        if (row < Mo) {
#pragma unroll
          for (int c = 0; c < CH / 32; ++c) {
            const int col0 = n0 + chalf * CH + c * 32;
            if (col0 >= No) continue;
            float* dst = C + (size_t)row * ldc + col0;
            const int nv = min(32, No - col0);
            float v[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = alpha * acc[c * 32 + j];
            const bool full32 = (nv == 32);
            if (ep.mode == EPI_FWD) {
              float bv[32];
              if (ep.bias && full32 && ((reinterpret_cast<uintptr_t>(ep.bias + col0) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 b4 = __ldg(reinterpret_cast<const float4*>(ep.bias + col0 + j));
                  bv[j] = b4.x; bv[j + 1] = b4.y; bv[j + 2] = b4.z; bv[j + 3] = b4.w;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) bv[j] = (ep.bias && j < nv) ? __ldg(ep.bias + col0 + j) : 0.f;
              }
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                float y = v[j] + bv[j];
                if (ep.act == GCBF_ACT_RELU) y = fmaxf(y, 0.f);
                else if (ep.act == GCBF_ACT_TANH) y = tanhf(y);
                v[j] = y;
              }
            } else if (ep.mode == EPI_DGRAD && ep.relu_src) {
              const float* ms = ep.relu_src + (size_t)row * ep.ld_relu + col0;
              if (full32 && ((reinterpret_cast<uintptr_t>(ms) & 15) == 0)) {
#pragma unroll
                for (int j = 0; j < 32; j += 4) {
                  const float4 m4 = __ldg(reinterpret_cast<const float4*>(ms + j));
                  v[j] = m4.x > 0.f ? v[j] : 0.f; v[j + 1] = m4.y > 0.f ? v[j + 1] : 0.f;
                  v[j + 2] = m4.z > 0.f ? v[j + 2] : 0.f; v[j + 3] = m4.w > 0.f ? v[j + 3] : 0.f;
                }
              } else {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                  if (j < nv) v[j] = (__ldg(ms + j) > 0.f) ? v[j] : 0.f;
              }
            }
            if (ep.mode == EPI_WGRAD && ep.atomic) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) atomicAdd(dst + j, v[j]);
            } else if (ep.accumulate) {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) dst[j] += v[j];
            } else if (nv == 32 && ((reinterpret_cast<uintptr_t>(dst) & 15) == 0)) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (j < nv) dst[j] = v[j];
            }
          }
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, K::TMEM_COLS);
  }
}

constexpr int TRANS_STRIP = 16;   // column tiles (of 32) per block in the transposing split

template <bool TRANS>
__global__ void split_kernel(const float* __restrict__ src, int ld, int rows, int cols, float* __restrict__ dst_hi,
                             float* __restrict__ dst_lo, int R_pad, int C_pad, float* __restrict__ rowsum, int strip) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  if (TRANS) {
    // source is [cols(src rows)][rows(src cols)]: read coalesced along r (source columns).  One block walks a strip of
    // TRANS_STRIP column tiles so the fused row sums need one atomic per row per strip (same-address atomics from many
    // blocks serialise in L2).
    float rs[4] = {0.f, 0.f, 0.f, 0.f};
    for (int ct = 0; ct < strip; ++ct) {
      const int cb = c0 * strip + ct * 32;
      if (cb >= C_pad) break;
      __syncthreads();
      for (int i = threadIdx.y; i < 32; i += 8) {
        const int c = cb + i, r = r0 + threadIdx.x;
        tile[i][threadIdx.x] = (c < cols && r < rows) ? src[(size_t)c * ld + r] : 0.f;
      }
      __syncthreads();
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int i = threadIdx.y + 8 * k;
        const int r = r0 + i, c = cb + threadIdx.x;
        const float x = tile[threadIdx.x][i];
        if (r < R_pad && c < C_pad) {
          float hi, lo;
          split_tf32(x, hi, lo);
          dst_hi[(size_t)r * C_pad + c] = hi;
          dst_lo[(size_t)r * C_pad + c] = lo;
        }
        rs[k] += x;
      }
    }
    if (rowsum) {   // fused bias gradient: sum over the (transposed) columns = column sums of the source
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float s = warp_sum(rs[k]);
        const int r = r0 + threadIdx.y + 8 * k;
        if (threadIdx.x == 0 && r < rows) atomicAdd(rowsum + r, s);
      }
    }
  } else {
    for (int i = threadIdx.y; i < 32; i += 8) {
      const int r = r0 + i, c = c0 + threadIdx.x;
      if (r < R_pad && c < C_pad) {
        const float x = (r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.f;
        float hi, lo;
        split_tf32(x, hi, lo);
        dst_hi[(size_t)r * C_pad + c] = hi;
        dst_lo[(size_t)r * C_pad + c] = lo;
      }
    }
  }
}

// ---- host side -----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && p) fn = (EncodeTiledFn)p;
  }
  return fn;
}

static int make_map(CUtensorMap* map, const float* base, int rows_total, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) { set_error("cuTensorMapEncodeTiled entry point unavailable"); return GCBF_E_CUDA; }
  cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows_total};
  cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
  cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, BK == 32 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d) rows=%d cols=%d box_rows=%d", (int)r, rows_total, cols, box_rows); return GCBF_E_CUDA; }
  return GCBF_OK;
}

// process-wide (one process per GPU): autograd runs backward on its own thread, so this must not be thread_local
static float* g_ws = nullptr;
static size_t g_ws_bytes = 0;
static int g_dbg = 0;
static bool g_a_raw = true;   // in-kernel split of operand A (GCBF_TC_A_RAW=0 in the environment disables it)

static inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

struct Operand {
  const float* src; int ld; int rows; int cols; bool trans;   // logical [rows][cols] K-major after optional transpose
};

static size_t operand_bytes(int rows, int cols, int tile_rows) {
  return (size_t)2 * round_up(rows, tile_rows) * round_up(cols, BK) * 4;
}

static int prep_operand(const Operand& op, int tile_rows, float* dst, int* R_pad_out, int* C_pad_out, cudaStream_t st,
                        float* rowsum = nullptr) {
  const int R_pad = round_up(op.rows, tile_rows), C_pad = round_up(op.cols, BK);
  float* hi = dst;
  float* lo = dst + (size_t)R_pad * C_pad;
  const int strip = (op.trans && rowsum) ? TRANS_STRIP : 1;   // long strips only where they save same-address atomics
  dim3 grid(ceil_div(C_pad, 32 * strip), ceil_div(R_pad, 32)), block(32, 8);
  if (op.trans) split_kernel<true><<<grid, block, 0, st>>>(op.src, op.ld, op.rows, op.cols, hi, lo, R_pad, C_pad, rowsum, strip);
  else split_kernel<false><<<grid, block, 0, st>>>(op.src, op.ld, op.rows, op.cols, hi, lo, R_pad, C_pad, nullptr, 1);
  GCBF_LAUNCH_OK();
  *R_pad_out = R_pad; *C_pad_out = C_pad;
  return GCBF_OK;
}

template <int BN, bool A_RAW>
static int launch_tiles(const float* a_ptr, int a_rows, int a_cols, int a_ld, const float* b_scr, int RB, int Kp, float* C,
                        int ldc, int Mo, int No, int splits, const EpiParams& ep, cudaStream_t st) {
  using K = Cfg<BN>;
  CUtensorMap ma, mb;
  // A_RAW: the map covers the caller's [Mo, Kc] matrix (OOB -> 0); otherwise the [hi; lo] scratch of 2*RA padded rows
  if (int rc = make_map(&ma, a_ptr, A_RAW ? a_rows : 2 * a_rows, A_RAW ? a_cols : Kp, a_ld, BM)) return rc;
  if (int rc = make_map(&mb, b_scr, 2 * RB, Kp, Kp, BN)) return rc;
  static bool attr_set = false;
  if (!attr_set) {
    GCBF_CUDA_OK(cudaFuncSetAttribute(gemm_tc_kernel<BN, A_RAW>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES));
    attr_set = true;
  }
  const int tiles_m = ceil_div(Mo, BM), tiles_n = RB / BN;
  const int kblocks = Kp / BK;
  const int kps = ceil_div(kblocks, splits);
  const int nsplit = ceil_div(kblocks, kps);
  int dev = 0, sms = kNumSMs;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const int ctas = max(1, min(tiles_m * tiles_n, max(1, sms / nsplit)));
  dim3 grid(ctas, nsplit);
  gemm_tc_kernel<BN, A_RAW><<<grid, NUM_THREADS, K::SMEM_BYTES, st>>>(ma, mb, a_rows, RB, C, ldc, Mo, No, tiles_m, tiles_n, kps,
                                                                      kblocks, ep);
  GCBF_LAUNCH_OK();
  return GCBF_OK;
}

// D[Mo,No] = A * B^T with A logical [Mo][Kc], B logical [No][Kc]; chunked over Mo (fwd/dgrad) or Kc (wgrad)
static int run_gemm(Operand A, Operand B, float* C, int ldc, EpiParams ep, bool chunk_k, cudaStream_t st, float* a_rowsum = nullptr) {
  const int Mo = A.rows, No = B.rows, Kc = A.cols;
  const int BN = (No > 128) ? 256 : 128;
  if (!g_ws) { set_error("tcgen05 GEMM: no workspace registered (gcbf_set_gemm_workspace)"); return GCBF_E_INVALID; }
  if (!chunk_k && !A.trans && (A.ld & 3) == 0 && (reinterpret_cast<uintptr_t>(A.src) & 15) == 0 && g_a_raw) {
    // operand A straight from the caller's matrix, split into hi / lo inside the kernel: only B needs scratch
    const size_t need = operand_bytes(No, Kc, BN);
    if (need > g_ws_bytes) { set_error("tcgen05 GEMM: workspace too small (%zu > %zu)", need, g_ws_bytes); return GCBF_E_INVALID; }
    int RB, Kp;
    if (int rc = prep_operand(B, BN, g_ws, &RB, &Kp, st)) return rc;
    ep.dbg = g_dbg;
    return (BN == 256) ? launch_tiles<256, true>(A.src, Mo, Kc, A.ld, g_ws, RB, Kp, C, ldc, Mo, No, 1, ep, st)
                       : launch_tiles<128, true>(A.src, Mo, Kc, A.ld, g_ws, RB, Kp, C, ldc, Mo, No, 1, ep, st);
  }
  if (!chunk_k) {
    for (int m0 = 0; m0 < Mo; m0 += MAX_CHUNK_ROWS) {
      const int mc = min(MAX_CHUNK_ROWS, Mo - m0);
      Operand a = A;
      a.rows = mc;
      a.src = A.trans ? A.src + m0 : A.src + (size_t)m0 * A.ld;
      const size_t need = operand_bytes(mc, Kc, BM) + operand_bytes(No, Kc, BN);
      if (need > g_ws_bytes) { set_error("tcgen05 GEMM: workspace too small (%zu > %zu)", need, g_ws_bytes); return GCBF_E_INVALID; }
      int RA, RB, Kp, Kp2;
      float* a_scr = g_ws;
      if (int rc = prep_operand(a, BM, a_scr, &RA, &Kp, st)) return rc;
      float* b_scr = g_ws + (size_t)2 * RA * Kp;
      if (int rc = prep_operand(B, BN, b_scr, &RB, &Kp2, st)) return rc;
      EpiParams e = ep;
      if (e.relu_src) e.relu_src = ep.relu_src + (size_t)m0 * ep.ld_relu;
      int rc = (BN == 256) ? launch_tiles<256, false>(a_scr, RA, Kp, Kp, b_scr, RB, Kp, C + (size_t)m0 * ldc, ldc, mc, No, 1, e, st)
                           : launch_tiles<128, false>(a_scr, RA, Kp, Kp, b_scr, RB, Kp, C + (size_t)m0 * ldc, ldc, mc, No, 1, e, st);
      if (rc) return rc;
    }
    return GCBF_OK;
  }
  // weight gradient: reduction dim (Kc = batch rows) is the big one -> chunk it and accumulate
  bool first = true;
  for (int k0 = 0; k0 < Kc; k0 += MAX_CHUNK_ROWS) {
    const int kc = min(MAX_CHUNK_ROWS, Kc - k0);
    Operand a = A, b = B;
    a.cols = kc; b.cols = kc;
    a.src = A.trans ? A.src + (size_t)k0 * A.ld : A.src + k0;
    b.src = B.trans ? B.src + (size_t)k0 * B.ld : B.src + k0;
    const size_t need = operand_bytes(Mo, kc, BM) + operand_bytes(No, kc, BN);
    if (need > g_ws_bytes) { set_error("tcgen05 GEMM: workspace too small (%zu > %zu)", need, g_ws_bytes); return GCBF_E_INVALID; }
    int RA, RB, Kp, Kp2;
    float* a_scr = g_ws;
    if (int rc = prep_operand(a, BM, a_scr, &RA, &Kp, st, a_rowsum)) return rc;
    float* b_scr = g_ws + (size_t)2 * RA * Kp;
    if (int rc = prep_operand(b, BN, b_scr, &RB, &Kp2, st)) return rc;
    const int tiles = (RA / BM) * (RB / BN);
    int splits = 1;
    if (tiles < kNumSMs) splits = max(1, min(Kp / 256, kNumSMs / tiles));
    EpiParams e = ep;
    e.atomic = splits > 1;
    if (!first) e.accumulate = 1;
    if (e.atomic && !e.accumulate) GCBF_CUDA_OK(cudaMemset2DAsync(C, (size_t)ldc * 4, 0, (size_t)No * 4, Mo, st));
    int rc = (BN == 256) ? launch_tiles<256, false>(a_scr, RA, Kp, Kp, b_scr, RB, Kp, C, ldc, Mo, No, splits, e, st)
                         : launch_tiles<128, false>(a_scr, RA, Kp, Kp, b_scr, RB, Kp, C, ldc, Mo, No, splits, e, st);
    if (rc) return rc;
    first = false;
  }
  return GCBF_OK;
}

}  // namespace tc

// ---- entry points used by linear.cu ---------------------------------------------------------------------------
static bool big_enough(long long M, long long N, long long K) { return M >= 256 && N >= 96 && K >= 64 && M * N * K >= (1ll << 24); }

bool tc_fwd_supported(int ldx, int ldw, int ldy, int M, int N, int K, bool forced) { return tc::g_ws && (forced || big_enough(M, N, K)); }
bool tc_dgrad_supported(int lddz, int ldw, int lddx, int M, int N, int K, bool forced) { return tc::g_ws && (forced || big_enough(M, K, N)); }
bool tc_wgrad_supported(int lddz, int ldx, int lddw, int M, int N, int K, bool forced) { return tc::g_ws && (forced || (N >= 96 && K >= 64 && M >= 256 && (long long)M * N * K >= (1ll << 24))); }

int launch_tc_fwd(const float* X, int ldx, const float* W, int ldw, const float* bias, const float* inv_sigma, float* Y,
                  int ldy, int M, int N, int K, int act, cudaStream_t st) {
  tc::EpiParams ep{};
  ep.mode = tc::EPI_FWD; ep.alpha = inv_sigma; ep.bias = bias; ep.act = act;
  return tc::run_gemm({X, ldx, M, K, false}, {W, ldw, N, K, false}, Y, ldy, ep, false, st);
}

int launch_tc_dgrad(const float* dZ, int lddz, const float* W, int ldw, const float* inv_sigma, const float* relu_src,
                    int ld_relu, float* dX, int lddx, int M, int N, int K, int accumulate, cudaStream_t st) {
  tc::EpiParams ep{};
  ep.mode = tc::EPI_DGRAD; ep.alpha = inv_sigma; ep.relu_src = relu_src; ep.ld_relu = ld_relu; ep.accumulate = accumulate;
  // dX[M,K] = dZ[M,N] * W[N,K]: A = dZ (contraction over N, contiguous), B = W^T  ([K][N])
  return tc::run_gemm({dZ, lddz, M, N, false}, {W, ldw, K, N, true}, dX, lddx, ep, false, st);
}

int launch_colsum(const float* dZ, int ld, int M, int N, float* db, int accumulate, cudaStream_t st);

int launch_tc_wgrad(const float* dZ, int lddz, const float* X, int ldx, const float* inv_sigma, float* dW, int lddw,
                    float* db, int M, int N, int K, int accumulate, cudaStream_t st) {
  tc::EpiParams ep{};
  ep.mode = tc::EPI_WGRAD; ep.alpha = inv_sigma; ep.accumulate = accumulate;
  // dW[N,K] = dZ^T[N,M] * X[M,K]: A = dZ^T ([N][M]), B = X^T ([K][M]); contraction over the M batch rows
  // the bias gradient (column sums of dZ) is fused into the transposing split of dZ
  if (db && !accumulate) GCBF_CUDA_OK(cudaMemsetAsync(db, 0, (size_t)N * 4, st));
  if (int rc = tc::run_gemm({dZ, lddz, N, M, true}, {X, ldx, K, M, true}, dW, lddw, ep, true, st, db)) return rc;
  return GCBF_OK;
}

}  // namespace gcbf

extern "C" int gcbf_set_gemm_workspace(void* ptr, size_t bytes) {
  const char* e = getenv("GCBF_TC_A_RAW");
  gcbf::tc::g_a_raw = !(e && e[0] == '0');
  const char* d = getenv("GCBF_TC_DBG");
  gcbf::tc::g_dbg = d ? atoi(d) : 0;
  gcbf::tc::g_ws = reinterpret_cast<float*>(ptr);
  gcbf::tc::g_ws_bytes = ptr ? bytes : 0;
  return GCBF_OK;
}

extern "C" size_t gcbf_gemm_workspace_bytes(int M, int N, int K) {
  // worst case over forward / data-grad / weight-grad of a [M,K] x [N,K] layer, with the launcher's chunking
  using namespace gcbf::tc;
  const int mc = M < MAX_CHUNK_ROWS ? M : MAX_CHUNK_ROWS;
  size_t fwd = operand_bytes(mc, K, BM) + operand_bytes(N, K, 256);
  size_t dgr = operand_bytes(mc, N, BM) + operand_bytes(K, N, 256);
  size_t wgr = operand_bytes(N, mc, BM) + operand_bytes(K, mc, 256);
  size_t m = fwd > dgr ? fwd : dgr;
  return (m > wgr ? m : wgr) + 4096;
}
