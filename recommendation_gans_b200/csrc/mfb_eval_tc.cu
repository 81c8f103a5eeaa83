// Tensor-core full-catalog scoring for evaluation (sm_100a: TMA + tcgen05.mma + TMEM).
//
// Replaces the score computation inside spotlight/evaluation.py:155-180 (model.predict(user) for every
// user = users x items x D flops) for the top-k metrics.  The user x item score matrix is never
// written to HBM:
//
//   1. k_tc_convert      fp16 copies of the user/item embedding rows (+ L2 norms for the error bound);
//      k_tc_xk_items / k_tc_xk_users: the 128 x 16 fp16 operands of the extra MMA step (item-bias pieces, threshold pieces)
//   2. k_tc_gemm<MAX>    scores of a SAMPLE of item tiles; every epilogue thread keeps the 24 largest maxima of its
//                        user's 32-item groups (groups holding a train item dropped)  ->  k_tc_threshold_merge: a
//                        per-user lower bound of the k-th best score (k > 24: group maxima through HBM + k_tc_threshold*)
//   3. k_tc_gemm<COLLECT> all item tiles; epilogue appends every (user, item) whose approximate score
//                        reaches the bound to the user's candidate list (a guaranteed superset of the top-k)
//   4. k_tc_rescore      exact fp32 re-scoring of the candidates (same sequential-FMA definition as
//                        k_topk_exact), train mask, top-k (ties -> lower item id); users whose list
//                        overflowed or cannot be certified join a list that k_topk_exact re-does (count on the device).
//
// GEMM mapping: A = 256 users of the CTA (two M = 128 blocks = the 128 TMEM lanes, twice; resident in shared memory),
// B = a tile of 128 items (N -> 128 TMEM columns per block), K = D (+ 16: the extra step that initialises the
// accumulator with item bias - user threshold), fp16 inputs, fp32 accumulation in TMEM.  An epilogue thread owns one
// USER (one TMEM lane).
// Warp roles (576 threads): warp 0 = TMA producer, warp 1 = MMA issuer / TMEM owner (one thread each, chosen with
// elect.sync), warps 2-17 = epilogue: (TMEM lane quarter) x (user block) x (tile parity).  Pipelines: smem full/empty
// (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue, four accumulator slots of 128 columns).
#include <cuda.h>
#include <cuda_fp16.h>
#include <math.h>
#include <stdlib.h>

#include "mfb_internal.cuh"

namespace {

constexpr int TC_M = 128;          // items per tile (the MMA's N: 128 TMEM columns per user block)
constexpr int TC_N = 256;          // users per CTA (two MMA M-blocks of 128 = the 128 TMEM lanes, twice)
#ifndef MFB_TC_STAGES
#define MFB_TC_STAGES 2
#endif
constexpr int TC_STAGES = MFB_TC_STAGES;   // item tiles in flight in shared memory (32 KB each at K = 128)
constexpr int TC_SROW = 36;        // row stride (floats) of an epilogue warp's score scratch: conflict-free 16-byte stores
#ifndef MFB_TC_SCR_ROWS
#define MFB_TC_SCR_ROWS 32
#endif
constexpr int TC_SCR_ROWS = MFB_TC_SCR_ROWS;   // scratch rows per epilogue warp: lanes with a hit in the chunk take one each
                                               // (32 = a private row per lane, no rounds)
#ifndef MFB_TC_PRESTORE_UNROLL
#define MFB_TC_PRESTORE_UNROLL 1
#endif
constexpr int TC_PRESTORE_UNROLL = MFB_TC_PRESTORE_UNROLL;   // 32-column chunks of the bias pre-store in flight
constexpr int TC_EPI_WARPS = 16;   // (TMEM lane quarter) x (user block) x (tile parity)
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;
constexpr int TC_KATOM = 64;       // 16-bit elements per 128-byte swizzle atom
constexpr int MODE_DUMP = 0, MODE_MAX = 1, MODE_COLLECT = 2;
constexpr int RS_MAXC = 512;       // candidate records kept per user (all sub-lists together)
constexpr int TC_CNT_STRIDE = 8;   // candidate counters per user: 2 tile parities x up to TC_MAX_SPLIT item-tile splits
constexpr int TC_MAX_SPLIT = 4;    // item-tile splits per user block (grid.y) when the user blocks alone cannot fill the SMs
constexpr float MASKED_SCORE_TC = -3.402823466e38f;

// ---------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
#ifdef MFB_TC_EPI_SLEEP
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
    __nanosleep(MFB_TC_EPI_SLEEP);
  }
#else
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
#endif
}
// same, for the single-thread producer / MMA roles: back off between polls so the spinning thread does not
// take issue slots from the epilogue warps that share its scheduler
__device__ __forceinline__ void mbar_wait_backoff(uint64_t *bar, uint32_t parity) {
  for (;;) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (done) break;
#ifndef MFB_TC_MMA_SPIN
    __nanosleep(64);
#endif
  }
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// 1-D bulk copy global -> shared (16-byte granularity), bytes counted on an mbarrier
__device__ __forceinline__ void bulk_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
// One thread of the (converged) warp.  With elect.sync the compiler knows that a single thread runs the region and emits
// bare UTCHMMA / UTMALDG instructions; under a plain `lane == 0` test it wraps every one of them in an ELECT / BRA.U.ANY
// loop over the active threads, and the issuing thread -- not the tensor pipe -- sets the pace (measured: ~105 instead of
// 64-71 cycles per 128x128x16 MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
#define TC_LD32(r, taddr)                                                                                            \
  asm volatile(                                                                                                      \
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                      \
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, "   \
      "%23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                                                         \
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),  \
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),       \
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),      \
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])                    \
      : "r"(taddr)                                                                                                   \
      : "memory")

#define TC_ST32(r, taddr)                                                                                              \
  asm volatile(                                                                                                        \
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                                  \
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, "  \
      "%24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),                                                    \
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),  \
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),     \
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),     \
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])                                                                   \
      : "memory")

// K-major, 128-byte-swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
// start address >> 4 in [0,14), leading byte offset >> 4 in [16,30) (unused for swizzled K-major: 1),
// stride byte offset >> 4 in [32,46) (1024 B between 8-row groups), version 1 in [46,48), SWIZZLE_128B = 2 in [61,64).
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// K-major operand WITHOUT swizzle: core matrices of 8 rows x 16 bytes; leading byte offset = distance between the core
// matrices adjacent in K (128 B), stride byte offset = distance between 8-row groups (256 B) -- the layout of xk_off.
__device__ __forceinline__ uint64_t umma_desc_noswz(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
  d |= (uint64_t)(128 >> 4) << 16;
  d |= (uint64_t)(256 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
// byte offset of element (row, k) of a 128 x 16 fp16 operand in that layout (4 KB per operand)
__host__ __device__ __forceinline__ int xk_off(int row, int k) {
  return (row >> 3) * 256 + (k >> 3) * 128 + (row & 7) * 16 + (k & 7) * 2;
}
constexpr int XK_BYTES = 4096;

// 256-bit read-only global load (32-byte aligned address)
__device__ __forceinline__ void ld_global_nc_v8(const float *p, float (&r)[8]) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(r[0]), "=f"(r[1]), "=f"(r[2]), "=f"(r[3]), "=f"(r[4]), "=f"(r[5]), "=f"(r[6]), "=f"(r[7])
               : "l"(p));
}

// order-preserving float -> int32 map (for integer redux max)
__device__ __forceinline__ int float_to_ordered(float f) {
  int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7fffffff);
}
__device__ __forceinline__ float ordered_to_float(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7fffffff)); }

// Item layout of the GEMM's A operand.  Item ids correlate with popularity in real catalogs (and in Zipf-drawn
// synthetic ones), which would put all of a user's best items into a handful of 32-item groups and make the
// sampled group maxima useless as a bound.  Items are therefore dealt round-robin over the T tiles and, inside
// a tile, spread over the four 32-lane quarters:  position = (item % T) * 128 + ((item / T) * 37 & 127).
// 37 * 45 = 1 (mod 128), so  item = ((slot * 45) & 127) * T + tile.  Positions whose item >= num_items are padding.
__device__ __forceinline__ int tc_item_of(int tile, int slot, int T) { return ((slot * 45) & 127) * T + tile; }
// magic = ceil(2^32 / T) gives item / T = umulhi(item, magic) exactly while item * T < 2^32 (the host passes 0 otherwise)
__device__ __forceinline__ int tc_pos_of(int item, int T, uint32_t magic) {
  const int j = magic ? (int)__umulhi((uint32_t)item, magic) : item / T;
  return (item - j * T) * TC_M + ((j * 37) & 127);
}

// ---------------------------------------------------------------------------------------------
// 1. fp32 -> fp16 row conversion (+ row norms).  rows_out >= rows: padding rows are zero.
//    `ids` (may be null) selects which source rows to convert (the evaluated users, in list order).
//    perm_T > 0: destination row r is a tile position, the source row is tc_item_of(r); the item biases are
//    copied into position order too.
// ---------------------------------------------------------------------------------------------
__global__ void k_tc_convert(const float *__restrict__ src, const long long *__restrict__ ids, int rows, int rows_out,
                             int D, int Dp, __half *__restrict__ dst, float *__restrict__ norm, float norm_scale,
                             float norm_offset, int perm_T, const float *__restrict__ bias_src,
                             float *__restrict__ bias_dst, float *__restrict__ norm_by_src,
                             int *__restrict__ overflow) {
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (r >= rows_out) return;
  float ss = 0.f;
  long long srow;
  if (perm_T > 0) {
    const int it = tc_item_of(r >> 7, r & 127, perm_T);
    srow = (it < rows) ? it : -1;
    // padding positions get a hugely negative bias: their scores can neither be a maximum nor reach a threshold
    if (lane == 0 && bias_dst) bias_dst[r] = (srow >= 0) ? bias_src[srow] : MASKED_SCORE_TC;
  } else {
    srow = (r < rows) ? (ids ? ids[r] : r) : -1;
  }
  bool big = false;
  for (int d = lane; d < Dp; d += 32) {     // Dp = D rounded up to the GEMM's K granularity: zero columns change no score
    float x = (srow >= 0 && d < D) ? src[srow * D + d] : 0.f;
    dst[(long long)r * Dp + d] = __float2half_rn(x);
    big |= !(fabsf(x) <= 60000.0f);    // beyond the fp16 range (or NaN): the error model below does not hold
    ss = fmaf(x, x, ss);
  }
  if (big) atomicExch(overflow, 1);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
  // the stored value is an UPPER bound of scale * |row| + offset: sqrtf and the fp32 sum of squares are rounded
  const float nrm = norm_scale * (sqrtf(ss) * 1.0001f) + norm_offset;
  if (lane == 0 && norm) norm[r] = nrm;
  if (lane == 0 && norm_by_src && srow >= 0) norm_by_src[srow] = nrm;
}

// per item tile: the largest (pre-scaled) row norm, one warp per tile
__global__ void k_tc_tile_maxnorm(const float *__restrict__ norm_pos, int tiles, float *__restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int t = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (t >= tiles) return;
  float m = 0.f;
  for (int j = lane; j < TC_M; j += 32) m = fmaxf(m, norm_pos[(long long)t * TC_M + j]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) out[t] = m;
}

// ---------------------------------------------------------------------------------------------
// 2/3. the GEMM kernel.  D[user][item] = sum_k U[user][k] * V[item][k]: the CTA's 256 users are the M side (two
// 128-row blocks = the 128 TMEM lanes, twice), a tile of 128 items is the N side (128 TMEM columns per block).
// An epilogue thread therefore owns ONE USER: its threshold, error radius and candidate counter are registers,
// the reduction over items is a chain of 3-input maxima inside the thread, and appends need no atomics.
// ---------------------------------------------------------------------------------------------
struct TcArgs {
  int num_items, n_users;       // valid rows of the item / user operand
  int D;                        // multiple of 64, <= 256
  int tile_begin, tile_step, n_tiles;   // item tiles processed: tile_begin + i*tile_step, i < n_tiles
  int total_tiles;              // T of the item layout (all tiles of the catalog)
  const float *item_bias;       // [items_pad]  item biases in position order; padding positions hold MASKED_SCORE_TC
  const float *tile_nmax;       // [total_tiles] max over the tile's items of err_coeff * |v|, where
                                //              |fp16-GEMM score - fp32 score| <= err_coeff * |u| * |v|
  const float *user_norm;       // [n_users_pad] L2 norm of each evaluated user's row
  // MODE_MAX: gmax[(i*4 + column group) * n_users_pad + user]  (ordered-int encoded)
  int *gmax;
  // MODE_MAX with TOPK > 0: the epilogue threads keep the TOPK largest clean group maxima themselves and write
  // toplists[((split * 2 + parity) * TOPK + j) * n_users_pad + user] (sorted, ordered-int encoded) instead of gmax;
  // dirty_groups[user][8]: sampled groups that hold a train item of the user (null: none)
  int *toplists;
  const uint32_t *dirty_groups;
  // XK kernels (bias and threshold enter through one extra K = 16 MMA step instead of a pre-store / a subtraction):
  // ximg[user block of 128][4 KB], yimg[item tile][4 KB]: the 128 x 16 fp16 operands in the no-swizzle K-major core
  // matrix layout (xk_off); rad_extra[user]: what the accumulation of the extra terms adds to the error radius
  const uint8_t *ximg;
  const uint8_t *yimg;
  const float *rad_extra;
  int n_users_pad;
  // MODE_COLLECT
  const float *thr;             // [n_users_pad] collection threshold per user
  int block_base;               // first user block of this launch (the grid covers a RANGE of user blocks)
  int2 *cand;                   // [n_users_pad][RS_MAXC]: sub-list (split * 2 + parity) of a user at offset (split * 2 + parity) * cap2: (item id, fp16-GEMM score incl. item bias, as float bits);
                                // the two sub-lists of a user belong to the two epilogue warps that share its columns
  int *cand_cnt;                // [n_users_pad][2]
  int cap2;
  // MODE_DUMP
  float *dump;                  // [num_items][n_users_pad]
  // train mask (null = no mask): a dense bitmap in exactly the order the epilogue threads consume it,
  // mask_bits[((cta * total_tiles + tile) * 16 + (q + 4*(ub + 2*ch))) * 32 + lane] = the 64 item-column bits of that
  // thread's user in that tile (x: columns 0..31 of the warp's half, y: 32..63); built by k_tc_mask_bitmap
  const uint2 *mask_bits;
  long long *timing;   // MFB_TC_TIMING builds: per-role cycle counters of two CTAs
  int dbg;   // experiment switches: 1 = skip score processing, 2 = skip appends, 4 = skip mask build, 8 = skip bias pre-store
};

// Pipelines: a ring of TC_STAGES item tiles in shared memory (TMA -> MMA) and FOUR accumulator slots of 128 columns in
// tensor memory (MMA -> epilogue): slot = (tile parity p) * 2 + (user block ub).  The 16 epilogue warps are split the
// same way: warp (lane quarter q, ub, p) drains all 128 item columns of block ub in the tiles i = p, p + 2, ... -- always
// slot (p, ub) -- so a warp's per-tile bookkeeping (tile index arithmetic, mask and bias fetches, barrier hand-shakes,
// ~115 instructions) is paid once per 128 columns, a slot is handed to its four warps as soon as its own 8 MMAs have
// retired, and its next use waits for those four warps only.  (The epilogue is bound by its instruction issue rate:
// ~110 instructions per 32 scores per warp in COLLECT mode; the MMA stream needs ~1024 of the ~2000 cycles of a tile.)
// XK = true: the item biases and (COLLECT) the user's collection threshold are the product of one extra K = 16 MMA
// step that INITIALISES the accumulator (accumulate = 0): X[user] = (1, 2^-6, 2^-12, -t_hi, -t_lo, 0...) against
// Y[item] = (b_hi, b_mid, b_lo, 1, 1, 0...), fp16 pieces whose products are exact in fp32 (k_tc_xk_items / _users).
// No bias pre-store (the slot goes back to the MMA as soon as it is drained) and no per-score subtraction: a hit is a
// clear sign bit.
template <int MODE, bool SPLIT, int TOPK, bool XK>
__global__ void __launch_bounds__(TC_THREADS, 1)   // 18 warps (allocated as 20): 96 registers per thread
k_tc_gemm(const __grid_constant__ CUtensorMap map_items, const __grid_constant__ CUtensorMap map_users,
          const TcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  // carve shared memory (1024-byte aligned operand buffers for the 128B swizzle)
  // (offset arithmetic on the shared-window address keeps every derived pointer a known shared-memory pointer,
  // so the compiler emits LDS / STS / ATOMS instead of generic accesses)
  uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  const int katoms = a.D / TC_KATOM;
  const uint32_t u_bytes = (uint32_t)TC_N * 128u * katoms;       // users, resident: [katom][256 rows][128 B]
  const uint32_t v_bytes = (uint32_t)TC_M * 128u * katoms;       // one item stage:  [katom][128 rows][128 B]
  uint8_t *sU = smem;
  uint8_t *sV = sU + u_bytes;
  uint8_t *sX = sV + (size_t)TC_STAGES * v_bytes;                // XK: [2 user blocks][4 KB]
  uint8_t *sY = sX + (XK ? 2 * XK_BYTES : 0);                    // XK: [TC_STAGES][4 KB]
  uint8_t *tail = sY + (XK ? TC_STAGES * XK_BYTES : 0);
  uint64_t *full = reinterpret_cast<uint64_t *>(tail);           // [TC_STAGES]
  uint64_t *empty = full + TC_STAGES;                            // [TC_STAGES]
  uint64_t *tfull = empty + TC_STAGES;                           // [4]: accumulator slot = buffer * 2 + user block
  uint64_t *tempty = tfull + 4;                                  // [4]
  uint64_t *ufull = tempty + 4;                                  // [1]
  static_assert((2 * TC_STAGES + 9) * 8 <= 120, "barriers overlap the TMEM slot word");
  uint32_t *tmem_slot = reinterpret_cast<uint32_t *>(tail + 120);
  float *sc_s = reinterpret_cast<float *>(tail + 128);                       // [16 warps][TC_SCR_ROWS][TC_SROW]
  float *bias_s = sc_s + TC_EPI_WARPS * TC_SCR_ROWS * TC_SROW;               // [16 warps][128]: item biases, prefetched

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Item tiles can be split over gridDim.y CTAs per user block (small user shards would otherwise leave SMs idle):
  // CTA (x, y) takes the tiles y, y + S, y + 2S, ... of the launch's tile sequence and owns its own candidate sub-lists.
  const int S = SPLIT ? (int)gridDim.y : 1, split = SPLIT ? (int)blockIdx.y : 0;   // SPLIT=false folds to the launch's own tile sequence
  const int tb = a.tile_begin + split * a.tile_step, ts = a.tile_step * S, nt = (a.n_tiles - split + S - 1) / S;
  const int ublock = (int)blockIdx.x + a.block_base;
  const int u0 = ublock * TC_N;
  // every CTA streams the same item tiles out of L2: start each CTA at a different tile so that concurrently
  // running CTAs do not all hit the same L2 slices at the same moment
  const int tile_off = (int)(((long long)blockIdx.x * 37) % (nt > 0 ? nt : 1));
  auto logical = [&](int i) { int li = i + tile_off; return li >= nt ? li - nt : li; };

  if (threadIdx.x == 0) {
    for (int s = 0; s < TC_STAGES; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    for (int b = 0; b < 4; ++b) {
      mbar_init(tfull + b, 1);
      mbar_init(tempty + b, TC_EPI_WARPS / 4);   // one arrival per epilogue warp of the slot (4 lane quarters)
    }
    mbar_init(ufull, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // TMEM: all 512 columns = 2 buffers x 2 user blocks x 128 item columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===== TMA producer =====
    if (elect_one()) {
      mbar_expect_tx(ufull, u_bytes + (XK ? 2 * XK_BYTES : 0));
      for (int ka = 0; ka < katoms; ++ka)
        tma_load_2d(sU + (size_t)ka * TC_N * 128, &map_users, ufull, ka * TC_KATOM, u0);
      if (XK) bulk_load_1d(sX, a.ximg + (size_t)(u0 / 128) * XK_BYTES, 2 * XK_BYTES, ufull);
#ifdef MFB_TC_TIMING
      long long tw_empty = 0;
#endif
      for (int i = 0; i < nt; ++i) {
        const int s = i % TC_STAGES;
        const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
#ifdef MFB_TC_TIMING
        const long long c0 = clock64();
#endif
        mbar_wait_backoff(empty + s, ph ^ 1u);
#ifdef MFB_TC_TIMING
        tw_empty += clock64() - c0;
        if ((a.dbg & 32) && i >= TC_STAGES) { mbar_arrive(full + s); continue; }   // ablation: no TMA traffic after the first fills
#endif
        const int row0 = (tb + logical(i) * ts) * TC_M;
        mbar_expect_tx(full + s, v_bytes + (XK ? XK_BYTES : 0));
        for (int ka = 0; ka < katoms; ++ka)
          tma_load_2d(sV + (size_t)s * v_bytes + (size_t)ka * TC_M * 128, &map_items, full + s, ka * TC_KATOM, row0);
        if (XK) bulk_load_1d(sY + (size_t)s * XK_BYTES, a.yimg + (size_t)(tb + logical(i) * ts) * XK_BYTES, XK_BYTES, full + s);
      }
#ifdef MFB_TC_TIMING
      if (a.timing && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && blockIdx.y == 0)
        a.timing[(blockIdx.x ? 24 : 0) + 0] = tw_empty;
#endif
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    // instruction descriptor (cute::UMMA::InstrDescriptor): D = F32 (bits 4-5 = 1), A = B = F16 (bits 7-9, 10-12 = 0),
    // both K-major (bits 15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28.  M = 128 users, N = 128 items.
    const uint32_t idesc = (1u << 4) | ((uint32_t)(TC_M >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    if (elect_one()) {
      mbar_wait_backoff(ufull, 0);
      tc_fence_after();
#ifdef MFB_TC_TIMING
      long long tw_tempty = 0, tw_full = 0, t_issue = 0;
      const long long t_start = clock64();
      unsigned long long g_start;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_start));
#endif
      // one (item tile, user block) unit at a time: its accumulator slot is handed over as soon as ITS 8 MMAs retire, and
      // the next use of a slot waits only for the four warps that drain that slot
      for (int j = 0; j < 2 * nt; ++j) {
        const int i = j >> 1, ub = j & 1;
        const int s = i % TC_STAGES;
        const uint32_t ph = (uint32_t)(i / TC_STAGES) & 1u;
        const int slot = (i & 1) * 2 + ub;
        const uint32_t bph = (uint32_t)(i >> 1) & 1u;
#ifdef MFB_TC_TIMING
        const long long c0 = clock64();
#endif
        mbar_wait_backoff(tempty + slot, bph);  // slot drained AND the item biases pre-stored by the epilogue
#ifdef MFB_TC_TIMING
        const long long c1 = clock64();
#endif
        if (ub == 0) mbar_wait_backoff(full + s, ph);     // item tile landed
#ifdef MFB_TC_TIMING
        const long long c2 = clock64();
        tw_tempty += c1 - c0;
        tw_full += c2 - c1;
#endif
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)(slot * 128);
#ifdef MFB_TC_TIMING
        const int ka_n = (a.dbg & 64) ? (katoms + 1) / 2 : katoms;   // ablation: half of the MMAs
#else
        const int ka_n = katoms;
#endif
        if (XK)   // bias (and threshold) terms: initialises the slot
          tc_mma_f16(d_tmem, umma_desc_noswz(smem_u32(sX + (size_t)ub * XK_BYTES)),
                     umma_desc_noswz(smem_u32(sY + (size_t)s * XK_BYTES)), idesc, 0u);
        for (int ka = 0; ka < ka_n; ++ka) {
          const uint64_t udesc = umma_desc_sw128(smem_u32(sU + (size_t)ka * TC_N * 128 + (size_t)ub * 128 * 128));
          const uint64_t vdesc = umma_desc_sw128(smem_u32(sV + (size_t)s * v_bytes + (size_t)ka * TC_M * 128));
#pragma unroll
          for (int k = 0; k < TC_KATOM / 16; ++k)   // 16 halves = 32 bytes per MMA along K: +2 in the (>>4) address field
            tc_mma_f16(d_tmem, udesc + (uint64_t)(2 * k), vdesc + (uint64_t)(2 * k), idesc, 1u);
        }
        if (ub == 1) tc_commit(empty + s);    // smem stage reusable once both blocks' MMAs retire
        tc_commit(tfull + slot);              // this block's accumulators ready for the epilogue
#ifdef MFB_TC_TIMING
        t_issue += clock64() - c2;
#endif
      }
#ifdef MFB_TC_TIMING
      if (a.timing && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && blockIdx.y == 0) {
        long long *t = a.timing + (blockIdx.x ? 24 : 0);
        unsigned long long g_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g_end));
        t[1] = tw_tempty; t[2] = tw_full; t[3] = t_issue; t[4] = clock64() - t_start; t[5] = nt; t[16] = (long long)(g_end - g_start);
      }
#endif
    }
  } else {
    // ===== epilogue warps: TMEM lane quarter q = warp % 4 (32 users), user block ub, tile parity par =====
    const int q = warp & 3;
    const int e = (warp - 2) >> 2;                         // 0..3
    const int ub = e & 1, par = e >> 1;
    const int ucol = ub * 128 + q * 32 + lane;             // this thread's user, as a column of the CTA
    const int gu = u0 + ucol;
    const bool user_ok = gu < a.n_users;
    const uint32_t lane_addr = ((uint32_t)(q * 32)) << 16;
    const uint32_t col_base = (uint32_t)(par * 256 + ub * 128);   // buffer par, block ub
    const bool use_mask = (MODE == MODE_COLLECT) && a.mask_bits != nullptr;
    float *scratch = sc_s + (warp - 2) * TC_SCR_ROWS * TC_SROW;
    float *my_bias = bias_s + (warp - 2) * TC_M;
    // this thread's mask words of a tile: two 8-byte loads (item-column halves), fetched one tile ahead of their use
    const uint2 *my_mask = use_mask ? a.mask_bits + ((long long)ublock * a.total_tiles * 16 + (q + 4 * ub)) * 32 + lane
                                    : nullptr;
    auto load_mask = [&](int tile_idx, int half) {
      if (!use_mask || tile_idx >= nt || (a.dbg & 4)) return make_uint2(0u, 0u);
      const int tile_id = tb + logical(tile_idx) * ts;
      return __ldg(my_mask + ((long long)tile_id * 16 + 8 * half) * 32);
    };
    uint2 mw_next0 = load_mask(par, 0), mw_next1 = load_mask(par, 1);
    const float nu = (MODE != MODE_DUMP && user_ok) ? a.user_norm[gu] : 0.f;
    const float thr_u = (MODE == MODE_COLLECT && user_ok) ? a.thr[gu] : INFINITY;
    const float rx = (XK && MODE == MODE_MAX && user_ok && a.rad_extra) ? a.rad_extra[gu] : 0.f;
    int my_cnt = 0;
    int2 *my_cand = (MODE == MODE_COLLECT) ? a.cand + (long long)gu * RS_MAXC + (split * 2 + par) * a.cap2 : nullptr;

    // item biases of a tile's 128 columns (they differ per column, not per user): fetched into this warp's slot with
    // cp.async at the top of the tile loop, written into the accumulator's next use at the bottom
    auto prefetch_bias = [&](int tile_idx) {
      if (XK) return;
      if (tile_idx < nt) {
        const int tile_id = tb + logical(tile_idx) * ts;
        const float *src = a.item_bias + (long long)tile_id * TC_M + lane * 4;
        asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(my_bias + lane * 4)), "l"(src) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    auto prestore_bias = [&]() {
      if (!XK) asm volatile("cp.async.wait_group 0;" ::: "memory");
      __syncwarp();
      if (!XK && !(a.dbg & 16)) {
#pragma unroll TC_PRESTORE_UNROLL
        for (int c0 = 0; c0 < TC_M; c0 += 32) {
          uint32_t r[32];
#pragma unroll
          for (int c4 = 0; c4 < 8; ++c4) {
            const float4 bv = *reinterpret_cast<const float4 *>(my_bias + c0 + c4 * 4);
            r[c4 * 4 + 0] = __float_as_uint(bv.x);
            r[c4 * 4 + 1] = __float_as_uint(bv.y);
            r[c4 * 4 + 2] = __float_as_uint(bv.z);
            r[c4 * 4 + 3] = __float_as_uint(bv.w);
          }
          TC_ST32(r, tmem_base + lane_addr + col_base + (uint32_t)c0);
        }
        asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty + par * 2 + ub);
    };
    if (par < nt) {
      prefetch_bias(par);
      prestore_bias();
      __syncwarp();
    }
    // MODE_MAX, TOPK > 0: the TOPK largest group maxima this thread has seen (groups without a train item of its user)
    int top[TOPK > 0 ? TOPK : 1];
#pragma unroll
    for (int j = 0; j < (TOPK > 0 ? TOPK : 1); ++j) top[j] = INT_MIN;
#ifdef MFB_TC_TIMING
    long long tw_tfull = 0, t_ldw = 0, t_proc = 0, t_pre = 0;
    const long long te_start = clock64();
#endif
    for (int i = par; i < nt; i += 2) {
      const uint32_t bph = (uint32_t)(i >> 1) & 1u;
      const int li = logical(i);
      const int tile_id = tb + li * ts;
      // error radius of every score of this tile for this user: |u| * max over the tile of err_coeff * |v|
      const float rad = (MODE != MODE_DUMP) ? nu * __ldg(a.tile_nmax + tile_id) : 0.f;
      prefetch_bias(i + 2);   // consumed by the pre-store at the bottom of this iteration
      // MODE_MAX, TOPK > 0: the word of the user's dirty-group set that holds this tile's four groups (fetched before the
      // wait on the accumulators)
      uint32_t dirty_w = 0u;
      if (MODE == MODE_MAX && TOPK > 0 && a.dirty_groups != nullptr && user_ok) {
        const int g0 = (li * S + split) * 4;
        if (g0 < 256) dirty_w = __ldg(a.dirty_groups + (long long)gu * 8 + (g0 >> 5));
      }
      const uint2 mw0 = mw_next0, mw1 = mw_next1;
      mw_next0 = load_mask(i + 2, 0);
      mw_next1 = load_mask(i + 2, 1);
#ifdef MFB_TC_TIMING
      const long long e0 = clock64();
#endif
      mbar_wait(tfull + par * 2 + ub, bph);
      tc_fence_after();
#ifdef MFB_TC_TIMING
      const long long e1 = clock64();
      tw_tfull += e1 - e0;
#endif
#pragma unroll 1
      for (int cc0 = 0; cc0 < ((a.dbg & 1) ? 0 : TC_M); cc0 += 32) {
        uint32_t r[32];
#ifdef MFB_TC_TIMING
        const long long l0 = clock64();
#endif
        TC_LD32(r, tmem_base + lane_addr + col_base + (uint32_t)cc0);
        // bit c: item column cc0+c is a train item of my user
        const uint32_t mword = (cc0 & 64) ? ((cc0 & 32) ? mw1.y : mw1.x) : ((cc0 & 32) ? mw0.y : mw0.x);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#ifdef MFB_TC_TIMING
        t_ldw += clock64() - l0;
#endif
        const int slot0 = cc0;                                   // tile slot of r[0]
        if (MODE == MODE_DUMP) {
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int item = tc_item_of(tile_id, slot0 + c, a.total_tiles);
            if (item < a.num_items) a.dump[(long long)item * a.n_users_pad + gu] = __uint_as_float(r[c]);
          }
        } else if (MODE == MODE_MAX) {
          // maximum over the 32 items, inside the thread.  The train mask is not applied here: the threshold kernel
          // drops every group that contains one of the user's train items; the bound only needs k clean groups.
          float mx = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
#pragma unroll
          for (int c = 2; c < 32; c += 2) mx = fmaxf(mx, fmaxf(__uint_as_float(r[c]), __uint_as_float(r[c + 1])));
          mx -= rad + rx;   // approx - err <= exact: a certified lower bound (train items: see k_tc_threshold_small)
          const int gi = (li * S + split) * 4 + (cc0 >> 5);   // the group's index in the sample
          if (TOPK > 0) {
            // the thread keeps the TOPK largest itself: no [groups][users] round trip through HBM, no selection kernel
            // behind it (k_tc_threshold_merge only merges the few lists of a user).  Groups that hold a train item of
            // the user are dropped: their maximum may belong to that item.
            int v = float_to_ordered(mx);
            if (!user_ok || ((dirty_w >> (gi & 31)) & 1u)) v = INT_MIN;
            if (__any_sync(0xffffffffu, v > top[(TOPK > 0 ? TOPK : 1) - 1])) {
#pragma unroll
              for (int j = 0; j < (TOPK > 0 ? TOPK : 1); ++j) {
                const int hi = max(top[j], v);
                v = min(top[j], v);
                top[j] = hi;
              }
            }
          } else {
            a.gmax[(long long)gi * a.n_users_pad + gu] = float_to_ordered(mx);
          }
        } else {
          // margin = score - (threshold - radius) on the FMA pipe; its sign bit (1 = below) is funnel-shifted into
          // one of four byte accumulators, columns taken from high to low so that column c lands on bit c
          const float t = thr_u - rad;    // approx + err >= exact: collect everything whose upper bound reaches thr
          uint32_t wb[4] = {0u, 0u, 0u, 0u};
          if (XK) {   // the accumulator already holds score - t' (t' <= thr - every radius): a hit is a clear sign bit
#pragma unroll
            for (int c = 31; c >= 0; --c) wb[c >> 3] = __funnelshift_l(r[c], wb[c >> 3], 1);
          } else {
#pragma unroll
            for (int c = 31; c >= 0; --c)
              wb[c >> 3] = __funnelshift_l(__float_as_uint(__uint_as_float(r[c]) - t), wb[c >> 3], 1);
          }
          uint32_t hw = ~(wb[0] | (wb[1] << 8) | (wb[2] << 16) | (wb[3] << 24)) & ~mword;
          if (!(a.dbg & 2)) {
            // the records carry the GEMM score (k_tc_rescore uses it to discard most of the list before the exact
            // pass); a hit's column is only known at run time, so a thread with hits copies its 32 scores to a
            // scratch row (the few lanes with hits take a row each, in rounds if there are more than TC_SCR_ROWS)
            if (TC_SCR_ROWS == 32) {
              if (hw != 0u) {
                float *row = scratch + lane * TC_SROW;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                  *reinterpret_cast<uint4 *>(row + c4 * 4) = make_uint4(r[c4 * 4 + 0], r[c4 * 4 + 1], r[c4 * 4 + 2], r[c4 * 4 + 3]);
                while (hw) {
                  const int c = __ffs(hw) - 1;
                  hw &= hw - 1u;
                  if (my_cnt < a.cap2)
                    my_cand[my_cnt] = make_int2(tc_item_of(tile_id, slot0 + c, a.total_tiles), __float_as_int(row[c]));
                  ++my_cnt;
                }
              }
            }
            unsigned pend = (TC_SCR_ROWS == 32) ? 0u : __ballot_sync(0xffffffffu, hw != 0u);
            while (pend) {
              const int rank = __popc(pend & ((1u << lane) - 1u));
              if (hw != 0u && rank < TC_SCR_ROWS) {
                float *row = scratch + rank * TC_SROW;
#pragma unroll
                for (int c4 = 0; c4 < 8; ++c4)
                  *reinterpret_cast<uint4 *>(row + c4 * 4) = make_uint4(r[c4 * 4 + 0], r[c4 * 4 + 1], r[c4 * 4 + 2], r[c4 * 4 + 3]);
                while (hw) {
                  const int c = __ffs(hw) - 1;
                  hw &= hw - 1u;
                  if (my_cnt < a.cap2)
                    my_cand[my_cnt] = make_int2(tc_item_of(tile_id, slot0 + c, a.total_tiles), __float_as_int(row[c]));
                  ++my_cnt;
                }
              }
              pend = __ballot_sync(0xffffffffu, hw != 0u);
            }
          }
        }
      }
#ifdef MFB_TC_TIMING
      const long long e2 = clock64();
      t_proc += e2 - e1;
#endif
      // hand the accumulators back: pre-store the biases of the tile that will use them next
      if (i + 2 < nt) prestore_bias();
#ifdef MFB_TC_TIMING
      t_pre += clock64() - e2;
#endif
    }
#ifdef MFB_TC_TIMING
    if (a.timing && (warp == 2 || warp == 17) && lane == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && blockIdx.y == 0) {
      long long *t = a.timing + (blockIdx.x ? 24 : 0) + (warp == 2 ? 6 : 11);
      t[0] = tw_tfull; t[1] = t_ldw; t[2] = t_proc; t[3] = t_pre; t[4] = clock64() - te_start;
    }
#endif
    if (MODE == MODE_COLLECT && user_ok) a.cand_cnt[(long long)gu * TC_CNT_STRIDE + split * 2 + par] = my_cnt;
    if (MODE == MODE_MAX && TOPK > 0) {
#pragma unroll
      for (int j = 0; j < (TOPK > 0 ? TOPK : 1); ++j)
        a.toplists[((long long)(split * 2 + par) * TOPK + j) * a.n_users_pad + gu] = top[j];
    }
  }
  // teardown
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
  }
}

// ---------------------------------------------------------------------------------------------
// threshold from the sampled group maxima: the m-th largest group maximum is a lower bound of the m-th best
// item score; m = k + (user's train items) so that at least k of the counted items are unmasked.
// thr_collect = bound - 2*eps_u (eps_u bounds |fp16-GEMM score - fp32 score|), certified later by k_tc_rescore.
// ---------------------------------------------------------------------------------------------
// ---------------------------------------------------------------------------------------------
// train mask for the COLLECT epilogue: a dense bitmap, built on the device per mfb_topk call.
// Block (c, s) owns 32 users of CTA-group c -- the users of the epilogue warps (q, ub) = (s & 3, s >> 2) -- and all
// item tiles (in ranges of MB_TILES): it assembles the 512 bytes per tile that those users' two epilogue warps
// (item-column halves) will read, in shared memory, from the users' CSR rows, and streams them out.  It also marks,
// per user, the sampled 32-item groups that contain a train item (`dirty`, 8 words per user) for
// k_tc_threshold_small.
// ---------------------------------------------------------------------------------------------
constexpr int MB_TILES = 400;     // 200 KB of bit images per range
constexpr int MB_THREADS = 256;

__global__ void __launch_bounds__(MB_THREADS) k_tc_mask_bitmap(const long long *__restrict__ user_ids,
                                                              const long long *__restrict__ indptr,
                                                              const int *__restrict__ indices, int n_users,
                                                              int total_tiles, uint32_t magic, int sample_step,
                                                              uint32_t *__restrict__ mask_bits,
                                                              uint32_t *__restrict__ dirty) {
  extern __shared__ uint32_t mb_bits[];          // [tiles of the range][2 halves][32 users][2 words]
  __shared__ uint32_t dirty_s[32][8];
  const int c = blockIdx.x, s = blockIdx.y;
  const int q = s & 3, ub = s >> 2;
  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  for (int i = tid; i < 32 * 8; i += MB_THREADS) (&dirty_s[0][0])[i] = 0u;
  for (int t0 = 0; t0 < total_tiles; t0 += MB_TILES) {
    const int nt = min(MB_TILES, total_tiles - t0);
    __syncthreads();
    for (int i = tid; i < nt * 128; i += MB_THREADS) mb_bits[i] = 0u;
    __syncthreads();
    for (int ul = wid; ul < 32; ul += MB_THREADS / 32) {   // warp per user, lanes over the user's entries
      const int p = c * TC_N + ub * 128 + q * 32 + ul;
      if (p >= n_users) continue;
      const long long uid = user_ids[p];
      const long long lo = indptr[uid], hi = indptr[uid + 1];
      for (long long base = lo; base < hi; base += 128) {
        int it[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {                       // four loads in flight per lane
          const long long e = base + j * 32 + lane;
          it[j] = (e < hi) ? indices[e] : -1;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (it[j] < 0) continue;
          const int pos = tc_pos_of(it[j], total_tiles, magic);
          const int tile = pos >> 7, col = pos & 127;
          if (tile >= t0 && tile < t0 + nt)
            atomicOr(mb_bits + (((tile - t0) * 2 + (col >> 6)) * 32 + ul) * 2 + ((col >> 5) & 1), 1u << (col & 31));
          if (t0 == 0) {
            const int li = tile / sample_step;
            if (li * sample_step == tile) {
              const int g = li * 4 + (col >> 5);
              if (g < 256) atomicOr(&dirty_s[ul][g >> 5], 1u << (g & 31));
            }
          }
        }
      }
    }
    __syncthreads();
    // the 256 bytes of (tile, half) go to epilogue warp q + 4*(ub + 2*half) of that tile's 4 KB image
    const uint4 *src = reinterpret_cast<const uint4 *>(mb_bits);
    for (int i = tid; i < nt * 32; i += MB_THREADS) {
      const int run = i >> 4, k16 = i & 15;
      const int t = t0 + (run >> 1), half = run & 1;
      uint4 *dst = reinterpret_cast<uint4 *>(mask_bits + (((long long)c * total_tiles + t) * 16 + (q + 4 * (ub + 2 * half))) * 64);
      dst[k16] = src[i];
    }
  }
  __syncthreads();
  for (int i = tid; i < 32 * 8; i += MB_THREADS)
    dirty[((long long)c * TC_N + ub * 128 + q * 32) * 8 + i] = (&dirty_s[0][0])[i];
}

constexpr int TH_VPL = 8;   // group maxima per lane -> up to 256 groups per user

// m-th largest group maximum, one THREAD per user (m <= K): the K largest keys are kept sorted in registers
// by a branch-free insertion (one max/min pair per slot); the loads of consecutive users coalesce.
// TPU = 4 (small user shards, where one thread per user leaves the SMs latency-bound): four neighbouring lanes share a
// user, each takes every fourth batch of groups, and the four sorted lists are merged with two rounds of shuffles.
template <int K, int TPU>
__global__ void __launch_bounds__(128) k_tc_threshold_small(const int *__restrict__ gmax, int groups, int n_users,
                                                            int n_users_pad, int m, float *__restrict__ thr,
                                                            const uint32_t *__restrict__ dirty_g) {
  // The MAX pass does not look at the train mask: a sampled 32-item group that contains one of the user's train
  // items is dropped here instead (its maximum may belong to that item).  dirty_g[user][8]: bit g of the user's
  // set (k_tc_mask_bitmap); null = no train mask.
  const int u = (blockIdx.x * 128 + threadIdx.x) / TPU;
  const int part = threadIdx.x % TPU;
  if (u >= n_users) return;
  const unsigned act = __activemask();   // lanes of this warp that own a user (uniform control flow from here on)
  uint32_t dirty[8];
  if (dirty_g != nullptr) {
    const uint4 d0 = *reinterpret_cast<const uint4 *>(dirty_g + (long long)u * 8);
    const uint4 d1 = *reinterpret_cast<const uint4 *>(dirty_g + (long long)u * 8 + 4);
    dirty[0] = d0.x; dirty[1] = d0.y; dirty[2] = d0.z; dirty[3] = d0.w;
    dirty[4] = d1.x; dirty[5] = d1.y; dirty[6] = d1.z; dirty[7] = d1.w;
  } else {
#pragma unroll
    for (int w = 0; w < 8; ++w) dirty[w] = 0u;
  }
  int top[K];
#pragma unroll
  for (int j = 0; j < K; ++j) top[j] = INT_MIN;
  const int *col = gmax + u;
  constexpr int NB = 8;          // loads in flight per thread; the next batch is issued before this one is inserted
  constexpr int GSTEP = NB * TPU;
  int nxt[NB];
#pragma unroll
  for (int t = 0; t < NB; ++t) nxt[t] = (part * NB + t < groups) ? col[(long long)(part * NB + t) * n_users_pad] : INT_MIN;
  for (int g = part * NB; g < groups; g += GSTEP) {
    int v[NB];
    uint32_t dsel = 0u;           // dirty[g >> 5] without a dynamically indexed register array
#pragma unroll
    for (int w = 0; w < 8; ++w) dsel = ((g >> 5) == w) ? dirty[w] : dsel;
    const uint32_t dw = dsel >> (g & 31);   // NB divides 32: the batch sits in one word
#pragma unroll
    for (int t = 0; t < NB; ++t) v[t] = ((dw >> t) & 1u) ? INT_MIN : nxt[t];   // INT_MIN never enters the top list
#pragma unroll
    for (int t = 0; t < NB; ++t)
      nxt[t] = (g + GSTEP + t < groups) ? col[(long long)(g + GSTEP + t) * n_users_pad] : INT_MIN;
#pragma unroll
    for (int t = 0; t < NB; ++t) {
      // most values are below the K-th largest seen so far: skip the K-deep network unless some lane of the warp needs it
      if (!__any_sync(act, v[t] > top[K - 1])) continue;
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int hi = max(top[j], v[t]);
        v[t] = min(top[j], v[t]);
        top[j] = hi;
      }
    }
  }
  if (TPU > 1) {   // merge the partners' lists: after round d every lane holds the top K of 2d lanes' groups
#pragma unroll
    for (int d = 1; d < TPU; d <<= 1) {
      int pv[K];
#pragma unroll
      for (int j = 0; j < K; ++j) pv[j] = __shfl_xor_sync(act, top[j], d);
#pragma unroll
      for (int i = 0; i < K; ++i) {
        int x = pv[i];
#pragma unroll
        for (int j = 0; j < K; ++j) {
          const int hi = max(top[j], x);
          x = min(top[j], x);
          top[j] = hi;
        }
      }
    }
    if (part != 0) return;
  }
  int r = INT_MIN;
#pragma unroll
  for (int j = 0; j < K; ++j) r = (j == m - 1) ? top[j] : r;
  thr[u] = (m <= groups && r != INT_MIN) ? ordered_to_float(r) : -INFINITY;   // -inf: the user goes to the exact path
}

// The m-th largest over a user's `nlists` sorted lists of K keys (written by the MAX epilogue threads that share the
// user: tile parities x item-tile splits); one thread per user, loads of consecutive users coalesce.
template <int K>
__global__ void __launch_bounds__(128) k_tc_threshold_merge(const int *__restrict__ lists, int nlists, int n_users,
                                                            int n_users_pad, int m, float *__restrict__ thr) {
  const int u = blockIdx.x * 128 + threadIdx.x;
  if (u >= n_users) return;
  const unsigned act = __activemask();
  int top[K];
#pragma unroll
  for (int j = 0; j < K; ++j) top[j] = lists[(long long)j * n_users_pad + u];   // list 0 is sorted already
  for (int l = 1; l < nlists; ++l) {
    const int *col = lists + (long long)l * K * n_users_pad + u;
    int v[K];
#pragma unroll
    for (int j = 0; j < K; ++j) v[j] = col[(long long)j * n_users_pad];
#pragma unroll
    for (int i = 0; i < K; ++i) {
      // the list is sorted: once its i-th key is below every lane's K-th largest, so are the rest
      if (!__any_sync(act, v[i] > top[K - 1])) break;
      int x = v[i];
#pragma unroll
      for (int j = 0; j < K; ++j) {
        const int hi = max(top[j], x);
        x = min(top[j], x);
        top[j] = hi;
      }
    }
  }
  int r = INT_MIN;
#pragma unroll
  for (int j = 0; j < K; ++j) r = (j == m - 1) ? top[j] : r;
  thr[u] = (r != INT_MIN) ? ordered_to_float(r) : -INFINITY;   // -inf: the user goes to the exact path
}

__global__ void __launch_bounds__(128) k_tc_threshold(const int *__restrict__ gmax, int groups, int n_users,
                                                      int n_users_pad, int k, const long long *__restrict__ user_ids,
                                                      const long long *__restrict__ indptr,
                                                      const float *__restrict__ unorm, const float *__restrict__ vmax,
                                                      float *__restrict__ thr, float *__restrict__ eps_out,
                                                      int masked_in_gemm) {
  const int lane = threadIdx.x & 31;
  const int u = blockIdx.x * 4 + (threadIdx.x >> 5);
  if (u >= n_users) return;
  long long ntrain = 0;
  if (indptr && !masked_in_gemm) {   // group maxima already exclude train items when the GEMM applies the mask
    const long long uid = user_ids[u];
    ntrain = indptr[uid + 1] - indptr[uid];
  }
  // the group maxima are already lower bounds of exact scores (approx - err_coeff*|u|*|v|), so the m-th largest
  // is a certified lower bound of the m-th best exact score: no further margin is needed
  if (lane == 0) eps_out[u] = 0.f;
  (void)unorm;
  (void)vmax;
  long long m = (long long)k + ntrain;
  if (m > groups) {
    if (lane == 0) thr[u] = -INFINITY;   // cannot bound: the user goes to the exact path
    return;
  }
  // values as order-preserving unsigned keys; radix-select the m-th largest from the top bit down
  uint32_t v[TH_VPL];
#pragma unroll
  for (int j = 0; j < TH_VPL; ++j) {
    const int g = j * 32 + lane;
    v[j] = (g < groups) ? ((uint32_t)gmax[(long long)g * n_users_pad + u] ^ 0x80000000u) : 0u;   // 0 = below everything
  }
  uint32_t prefix = 0, mask = 0;
  int want = (int)m;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t test = prefix | (1u << bit);
    const uint32_t tmask = mask | (1u << bit);
    int c = 0;
#pragma unroll
    for (int j = 0; j < TH_VPL; ++j) c += ((v[j] & tmask) == test) ? 1 : 0;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (c >= want) {
      prefix = test;      // the m-th largest has this bit set
    } else {
      want -= c;          // skip the c values above
    }
    mask = tmask;
  }
  if (lane == 0) thr[u] = ordered_to_float((int)(prefix ^ 0x80000000u));
}

// ---------------------------------------------------------------------------------------------
// XK operands (see k_tc_gemm<..., XK = true>).  fp16 x fp16 products are exact in fp32, so
//   b  = b_hi + 2^-6 b_mid + 2^-12 b_lo   (three 11-bit pieces: residual <= 2^-33 |b|, far below one fp32 ulp)
//   t' = t_hi + t_lo                       (DEFINED as the sum of its two pieces, chosen <= the wanted threshold)
// enter the accumulator exactly; what the tensor core's fp32 accumulation of these larger terms may add to the
// error of a score is bounded by rad_extra (2^-16 of their magnitudes, see k_tc_xk_users, + the pieces' underflow).
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(TC_M) k_tc_xk_items(const float *__restrict__ bias_pos, int total_tiles,
                                                      uint8_t *__restrict__ yimg, float *__restrict__ bmax,
                                                      int *__restrict__ overflow) {
  const int tile = blockIdx.x, row = threadIdx.x;
  float b = bias_pos[(long long)tile * TC_M + row];
  __half y[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) y[k] = __float2half_rn(0.f);
  if (b == MASKED_SCORE_TC) {            // padding position: never a maximum, never a hit
    y[0] = __float2half_rn(-60000.0f);
  } else {
    if (!(fabsf(b) <= 60000.0f)) atomicExch(overflow, 1);
    const __half h0 = __float2half_rn(b);
    const float r1 = (b - __half2float(h0)) * 64.0f;            // exact
    const __half h1 = __float2half_rn(r1);
    const float r2 = (r1 - __half2float(h1)) * 64.0f;           // exact
    y[0] = h0;
    y[1] = h1;
    y[2] = __float2half_rn(r2);
    atomicMax(reinterpret_cast<int *>(bmax), __float_as_int(fabsf(b)));   // non-negative floats order as ints
  }
  y[3] = __float2half_rn(1.0f);
  y[4] = __float2half_rn(1.0f);
  uint8_t *dst = yimg + (size_t)tile * XK_BYTES;
  *reinterpret_cast<uint4 *>(dst + xk_off(row, 0)) = *reinterpret_cast<const uint4 *>(&y[0]);
  *reinterpret_cast<uint4 *>(dst + xk_off(row, 8)) = *reinterpret_cast<const uint4 *>(&y[8]);
}

// collect = 0: X = (1, 2^-6, 2^-12, 0, ...), rad_extra = what the bias terms add to the radius (MAX pass)
// collect = 1: X additionally carries -t' for t' <= thr - |u| max_tile(nmax) - rad_extra, t' and rad_extra written out
__global__ void __launch_bounds__(128) k_tc_xk_users(int n_users, int n_users_pad, int collect,
                                                     const float *__restrict__ thr, const float *__restrict__ unorm,
                                                     const float *__restrict__ tile_nmax, int total_tiles,
                                                     const float *__restrict__ bmax, uint8_t *__restrict__ ximg,
                                                     float *__restrict__ tprime, float *__restrict__ rad_extra,
                                                     int *__restrict__ overflow) {
  const int u = blockIdx.x * 128 + threadIdx.x;
  if (u >= n_users_pad) return;
  __half x[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = __float2half_rn(0.f);
  x[0] = __float2half_rn(1.0f);
  x[1] = __float2half_rn(0.015625f);          // 2^-6
  x[2] = __float2half_rn(0.000244140625f);    // 2^-12
  const float bm = *bmax;
  // 2^-16 of the extra terms' magnitude M = |b| + |t|.  The accumulator now sits at magnitude M while the D products of
  // the dot are added to it: if every addend of a K = 16 step is aligned to the largest exponent and TRUNCATED there with
  // g >= 2 guard bits (the behaviour measured on earlier tensor-core generations; the rounding is not specified), a step
  // loses at most 17 * 2^-(23+g) M and the 9 steps 2^-17.7 M.  2^-16 leaves a factor of three.  + 2^-23 absolute for
  // pieces that underflow in fp16 and for t' = t_hi + t_lo formed in fp32 by the re-score.
  float rx = 1.52587890625e-05f * bm + 1.1920929e-07f;
  float tp = 0.f;
  if (collect) {
    float t_raw = -INFINITY;
    if (u < n_users && thr[u] > -INFINITY) {
      float nmax = 0.f;
      for (int t = 0; t < total_tiles; ++t) nmax = fmaxf(nmax, __ldg(tile_nmax + t));
      t_raw = thr[u] - unorm[u] * nmax * 1.000001f;
    }
    if (t_raw > -INFINITY) {
      rx += 1.52587890625e-05f * fabsf(t_raw);
      const float t0 = t_raw - rx - 9.5367431640625e-07f * fabsf(t_raw);   // room for the two-piece representation
      if (!(fabsf(t0) <= 60000.0f)) atomicExch(overflow, 1);
      const __half th = __float2half_rn(t0);
      const __half tl = __float2half_rn(t0 - __half2float(th));
      tp = __half2float(th) + __half2float(tl);          // exact: two 11-bit pieces
      x[3] = __hneg(th);
      x[4] = __hneg(tl);
    } else {
      // no bound for this user (it is re-done by the exact kernel): a threshold nothing reaches
      x[3] = __float2half_rn(-60000.0f);
      tp = 60000.0f;
    }
    tprime[u] = tp;
  }
  rad_extra[u] = rx;
  uint8_t *dst = ximg + (size_t)(u >> 7) * XK_BYTES;
  const int row = u & 127;
  *reinterpret_cast<uint4 *>(dst + xk_off(row, 0)) = *reinterpret_cast<const uint4 *>(&x[0]);
  *reinterpret_cast<uint4 *>(dst + xk_off(row, 8)) = *reinterpret_cast<const uint4 *>(&x[8]);
}

// ---------------------------------------------------------------------------------------------
// 4. exact re-score of the candidates + mask + top-k.  One warp per user.
//    Exact score = sequential fp32 FMA over d = 0..D-1, then (+ user bias) + item bias: bit-identical to k_topk_exact.
// ---------------------------------------------------------------------------------------------
constexpr int RS_WARPS = 4;
#ifndef MFB_RS_INFLIGHT
#define MFB_RS_INFLIGHT 4
#endif
constexpr int RS_INFLIGHT = MFB_RS_INFLIGHT;   // 256-bit row loads in flight per lane in the exact re-score (D % (8 * RS_INFLIGHT) == 0 on the vector path)

// smem per warp: user row [D] | scores [RS_MAXC] (upper bounds, then exact) | lower bounds [RS_MAXC] | ids [RS_MAXC]
// (staging the candidate rows through shared memory with cp.async was measured 2x slower: the kernel is
// latency-bound and the extra shared memory cuts the resident warps)
template <int RS_MAXSUB>   // upper bound of nsub: 2 (no item-tile split) or 2 * TC_MAX_SPLIT
__global__ void __launch_bounds__(RS_WARPS * 32) k_tc_rescore(
    const long long *__restrict__ user_ids, int n_users, TableView users, TableView items, int D,
    const int2 *__restrict__ cand, const int *__restrict__ cand_cnt, int cap2, int nsub, const float *__restrict__ thr,
    const float *__restrict__ unorm, const float *__restrict__ item_norm,
    const long long *__restrict__ indptr, const int *__restrict__ indices, int k, int *__restrict__ out_ids,
    float *__restrict__ out_scores, int *__restrict__ redo_flag, int *__restrict__ surv_cnt, int check_mask,
    long long *__restrict__ redo_users, int *__restrict__ redo_pos, int *__restrict__ redo_cnt,
    const float *__restrict__ tprime, const float *__restrict__ rad_extra, const float *__restrict__ tile_nmax,
    int total_tiles, uint32_t tile_magic, int pos_base) {
  // (the launch covers the users [pos_base, pos_base + n_users) of the call: every per-user pointer is offset by the host)
  // (XK GEMM: the records hold score - t'[u], and the accumulation of the extra terms widens every bound by rad_extra[u])
  // A user that cannot be certified here joins the list the exact kernel re-does (order arbitrary: every entry names
  // its own output row).  redo_cnt[1] = the fp16-range overflow flag of k_tc_convert: no certificate holds at all.
  auto give_up = [&](int uu, long long uuid) {
    redo_flag[uu] = 1;
    const int at = atomicAdd(redo_cnt, 1);
    redo_users[at] = uuid;
    redo_pos[at] = pos_base + uu;
  };
  extern __shared__ __align__(16) float rs_smem[];
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int u = blockIdx.x * RS_WARPS + wib;
  float *urow = rs_smem + (size_t)wib * (D + 3 * RS_MAXC);
  float *sc = urow + D;
  float *lob = sc + RS_MAXC;
  int *ids = reinterpret_cast<int *>(lob + RS_MAXC);
  if (u >= n_users) return;
  const long long uid = user_ids[u];
  // the user's nsub sub-lists (item-column halves x item-tile splits): pre[j] = entries before sub-list j
  int pre[RS_MAXSUB + 1];
  bool over = false;
  pre[0] = 0;
#pragma unroll
  for (int j = 0; j < RS_MAXSUB; ++j) {
    const int cj = (j < nsub) ? cand_cnt[(long long)u * TC_CNT_STRIDE + j] : 0;
    over |= cj > cap2;
    pre[j + 1] = pre[j] + cj;
  }
  const int cnt = pre[RS_MAXSUB];
  if (over || cnt > RS_MAXC || cnt < k || !(thr[u] > -INFINITY) || redo_cnt[1] != 0) {   // overflow / no bound: exact path
    if (lane == 0) {
      give_up(u, uid);
      surv_cnt[u] = 0;
    }
    return;
  }
  const bool vec = (D % (8 * RS_INFLIGHT)) == 0;      // whole chunks: vector loads; otherwise (e.g. the CLI's default D = 50) scalar
  if (vec) {
    for (int d = lane * 4; d < D; d += 128) *reinterpret_cast<float4 *>(urow + d) = *reinterpret_cast<const float4 *>(users.p + uid * D + d);
  } else {
    for (int d = lane; d < D; d += 32) urow[d] = users.p[uid * D + d];
  }
  const float ub = users.bp[uid];
  long long tlo = 0, thi = 0;
  if (indptr && check_mask) {    // (the GEMM epilogue already dropped train items when it applied the mask itself)
    tlo = indptr[uid];
    thi = indptr[uid + 1];
  }
  // ---- 1. bounds of every listed item's exact score from its GEMM score: [s - e, s + e], e = err_coeff*|u|*|v|
  const float nu = unorm[u];
  const float tp = tprime ? tprime[u] : 0.f, rxu = rad_extra ? rad_extra[u] : 0.f;
  constexpr int KR = 8;                       // lists of up to 32*KR items keep their lower-bound keys in registers
  const bool small = cnt <= 32 * KR;
  uint32_t kreg[KR];
#pragma unroll
  for (int j = 0; j < KR; ++j) kreg[j] = 0u;  // 0 never matches a radix-select test
  for (int c0 = 0; c0 < cnt; c0 += 32 * KR) {
#pragma unroll
    for (int j = 0; j < KR; ++j) {
      const int c = c0 + j * 32 + lane;
      if (c < cnt) {
        int sub = 0;
#pragma unroll
        for (int j = 1; j < RS_MAXSUB; ++j) sub += (c >= pre[j]) ? 1 : 0;      // pre[] is non-decreasing
        sub = sub < nsub ? sub : nsub - 1;
        const int2 rec = cand[(long long)u * RS_MAXC + sub * cap2 + (c - pre[sub])];
        // err_coeff * |v|: the item's own value (a gather: one 32-byte sector per lane), or -- tile_nmax given -- the
        // maximum over the item's tile (item % T in the item layout), a table that stays in L1: a slightly wider, still
        // valid bound for a quarter of this kernel's L1 look-ups
        float e;
        if (tile_nmax != nullptr) {
          const int jq = tile_magic ? (int)__umulhi((uint32_t)rec.x, tile_magic) : rec.x / total_tiles;
          e = __ldg(tile_nmax + (rec.x - jq * total_tiles));
        } else {
          e = item_norm[rec.x];
        }
        const float sg = __int_as_float(rec.y) + tp;
        const float lo = fmaf(-e, nu, sg) - rxu;
        ids[c] = rec.x;
        sc[c] = fmaf(e, nu, sg) + rxu;
        lob[c] = lo;
        if (c0 == 0) kreg[j] = (uint32_t)float_to_ordered(lo) ^ 0x80000000u;
      }
    }
  }
  __syncwarp();
  // ---- 2. the k-th largest lower bound L: at least k listed items have exact score >= L, so an item whose upper
  // bound is below L is strictly below the k-th best and cannot be in the top-k.  Radix select from the top bit.
  int n_surv = cnt;
  if (!check_mask) {   // (with the mask applied here instead of in the GEMM, listed items may be train items: no filter)
    // Only the top 20 key bits are resolved: the result is the k-th largest key with its low bits cleared, a
    // slightly smaller -- still valid -- bound.  Lists of up to 256 items (the usual case) sit in registers.
    uint32_t prefix = 0, mask = 0;
    int want = k;
    const int nv = (cnt + 31) >> 5;           // key registers in use (warp-uniform)
    for (int bit = 31; bit >= 12; --bit) {
      const uint32_t test = prefix | (1u << bit), tmask = mask | (1u << bit);
      int c1 = 0;
      if (small) {
#pragma unroll
        for (int j = 0; j < KR; ++j)
          if (j < nv) c1 += ((kreg[j] & tmask) == test) ? 1 : 0;
      } else {
        for (int c = lane; c < cnt; c += 32)
          c1 += ((((uint32_t)float_to_ordered(lob[c]) ^ 0x80000000u) & tmask) == test) ? 1 : 0;
      }
      c1 = __reduce_add_sync(0xffffffffu, c1);
      if (c1 >= want) prefix = test; else want -= c1;
      mask = tmask;
    }
    const float L = ordered_to_float((int)(prefix ^ 0x80000000u));
    // compact the survivors (upper bound >= L) to the front, order preserved
    n_surv = 0;
    for (int base = 0; base < cnt; base += 32) {
      const int c = base + lane;
      const bool keep = c < cnt && sc[c] >= L;
      const int id = c < cnt ? ids[c] : 0;
      const unsigned bal = __ballot_sync(0xffffffffu, keep);
      __syncwarp();
      if (keep) ids[n_surv + __popc(bal & ((1u << lane) - 1u))] = id;
      n_surv += __popc(bal);
      __syncwarp();
    }
  }
  if (lane == 0) surv_cnt[u] = n_surv;
  // ---- 3. exact scores of the survivors
  // every item outside the list has exact <= approx + err < thr; k listed, unmasked items with exact >= thr
  // (guaranteed by construction: the sampled items behind the bound are listed) prove the top-k is inside the list
  const float certify = thr[u];
  int good = 0;
  for (int c = lane; c < n_surv; c += 32) {
    const int my_item = ids[c];
    const float *v = items.p + (long long)my_item * D;
    const float ib = items.bp[my_item];
    // exact score: sequential fp32 FMA over d = 0..D-1 (bit-identical to k_topk_exact); 8 row loads in flight
    // Every lane walks its own row, so one warp-wide load instruction touches 32 different cache lines: the kernel is
    // bound by the L1 tag rate (ncu: l1tex 82 % of peak), not by bytes.  256-bit loads (sm_100: LDG.256, one whole
    // 32-byte sector per lane) halve the number of look-ups per row.
    float acc = 0.f;
    if (!vec) {
      for (int d = 0; d < D; ++d) acc = fmaf(urow[d], __ldg(v + d), acc);
    } else
    for (int d0 = 0; d0 < D; d0 += 8 * RS_INFLIGHT) {
      float x[RS_INFLIGHT][8];
#pragma unroll
      for (int j = 0; j < RS_INFLIGHT; ++j) ld_global_nc_v8(v + d0 + 8 * j, x[j]);
#pragma unroll
      for (int j = 0; j < RS_INFLIGHT; ++j) {
        const float4 y0 = *reinterpret_cast<const float4 *>(urow + d0 + 8 * j);
        const float4 y1 = *reinterpret_cast<const float4 *>(urow + d0 + 8 * j + 4);
        acc = fmaf(y0.x, x[j][0], acc);
        acc = fmaf(y0.y, x[j][1], acc);
        acc = fmaf(y0.z, x[j][2], acc);
        acc = fmaf(y0.w, x[j][3], acc);
        acc = fmaf(y1.x, x[j][4], acc);
        acc = fmaf(y1.y, x[j][5], acc);
        acc = fmaf(y1.z, x[j][6], acc);
        acc = fmaf(y1.w, x[j][7], acc);
      }
    }
    float z = (acc + ub) + ib;
    const float zc = acc + ib;   // the GEMM's score space has no user bias (constant per user)
    long long lo = tlo, hi = thi;
    while (lo < hi) {            // train mask: binary search in the user's sorted CSR row
      long long mid = (lo + hi) >> 1;
      if (indices[mid] < my_item) lo = mid + 1; else hi = mid;
    }
    const bool masked = (lo < thi) && indices[lo] == my_item;
    if (masked) z = MASKED_SCORE_TC;
    else if (zc >= certify) ++good;
    sc[c] = z;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) good += __shfl_xor_sync(0xffffffffu, good, o);
  if (good < k) {   // the collected set is not certified to contain the exact top-k
    if (lane == 0) give_up(u, uid);
    return;
  }
  __syncwarp();
  // ---- 4. top-k by (score desc, id asc)
  if (n_surv <= 32) {
    // one survivor per lane: bitonic sort of 64-bit keys (ordered score : inverted id), largest first
    unsigned long long key = 0ull;
    if (lane < n_surv)
      key = ((unsigned long long)((uint32_t)float_to_ordered(sc[lane]) ^ 0x80000000u) << 32) |
            (unsigned long long)(uint32_t)(0x7fffffff - ids[lane]);
#pragma unroll
    for (int k2 = 2; k2 <= 32; k2 <<= 1) {
#pragma unroll
      for (int j = k2 >> 1; j > 0; j >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, key, j);
        const bool desc = (lane & k2) == 0;            // this block ends up largest-first (k2 == 32: everyone)
        const bool lower = (lane & j) == 0;            // this lane keeps the block's "first" element of the pair
        const bool take_max = (desc == lower);
        key = take_max ? (key > other ? key : other) : (key < other ? key : other);
      }
    }
    if (lane < k) {
      const float bv = ordered_to_float((int)((uint32_t)(key >> 32) ^ 0x80000000u));
      out_ids[(long long)u * k + lane] = 0x7fffffff - (int)(uint32_t)(key & 0xffffffffull);
      if (out_scores) out_scores[(long long)u * k + lane] = (bv == MASKED_SCORE_TC) ? 0.f : 1.0f / (1.0f + expf(-bv));
    }
    if (lane == 0) redo_flag[u] = 0;
    return;
  }
  // longer lists: k rounds of warp arg-max; winners are removed
  for (int r = 0; r < k; ++r) {
    float bv = -INFINITY;
    int bi = 0x7fffffff, bc = -1;
    for (int c = lane; c < n_surv; c += 32) {
      const float z = sc[c];
      const int id = ids[c];
      if (id >= 0 && (z > bv || (z == bv && id < bi))) {
        bv = z;
        bi = id;
        bc = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      const int oc = __shfl_xor_sync(0xffffffffu, bc, o);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
        bc = oc;
      }
    }
    if (lane == 0) {
      out_ids[(long long)u * k + r] = bi;
      if (out_scores) out_scores[(long long)u * k + r] = (bv == MASKED_SCORE_TC) ? 0.f : 1.0f / (1.0f + expf(-bv));
      if (bc >= 0) ids[bc] = -1;
    }
    __syncwarp();
  }
  if (lane == 0) redo_flag[u] = 0;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                        const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                        CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);

int make_tmap(CUtensorMap *map, void *base, int rows, int D, int box_rows) {
  static PFN_tmapEncodeTiled encode = nullptr;
  if (!encode) {
    void *fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      mfb_set_error("cuTensorMapEncodeTiled unavailable (%s)", cudaGetErrorString(e));
      return MFB_ERR_CUDA;
    }
    encode = (PFN_tmapEncodeTiled)fn;
  }
  cuuint64_t gdim[2] = {(cuuint64_t)D, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)D * sizeof(__half)};
  cuuint32_t box[2] = {(cuuint32_t)TC_KATOM, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = encode(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, gdim, gstride, box, estr,
                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mfb_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r);
    return MFB_ERR_CUDA;
  }
  return MFB_OK;
}

size_t tc_smem_bytes(int D, bool xk = false) {
  const int katoms = D / TC_KATOM;
  return 1024 + (size_t)TC_N * 128 * katoms + (size_t)TC_STAGES * TC_M * 128 * katoms + 128 +
         (xk ? (size_t)(2 + TC_STAGES) * XK_BYTES : 0) +
         (size_t)TC_EPI_WARPS * TC_SCR_ROWS * TC_SROW * 4 +
         (size_t)TC_EPI_WARPS * TC_M * 4;
}

template <int MODE, bool SPLIT, int TOPK, bool XK>
int launch_gemm_s(const CUtensorMap &mi, const CUtensorMap &mu, const TcArgs &a, dim3 grid, size_t smem, cudaStream_t st) {
  static size_t smem_set = 0;   // (one device per process: the attribute is set when the size first grows)
  if (smem > smem_set) {
    MFB_CUDA(cudaFuncSetAttribute(k_tc_gemm<MODE, SPLIT, TOPK, XK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    smem_set = smem;
  }
  k_tc_gemm<MODE, SPLIT, TOPK, XK><<<grid, TC_THREADS, smem, st>>>(mi, mu, a);
  MFB_KERNEL_CHECK();
  return MFB_OK;
}

// grid: the user blocks [a.block_base, a.block_base + n_blocks) x splits
template <int MODE, int TOPK = 0>
int launch_gemm(const CUtensorMap &mi, const CUtensorMap &mu, const TcArgs &a, int n_blocks, cudaStream_t st,
                int splits = 1) {
  const bool xk = a.ximg != nullptr;
  const size_t smem = tc_smem_bytes(a.D, xk);
  const dim3 grid(n_blocks, splits);
  if (xk) {
    if (splits > 1) return launch_gemm_s<MODE, true, TOPK, true>(mi, mu, a, grid, smem, st);
    return launch_gemm_s<MODE, false, TOPK, true>(mi, mu, a, grid, smem, st);
  }
  if (splits > 1) return launch_gemm_s<MODE, true, TOPK, false>(mi, mu, a, grid, smem, st);
  return launch_gemm_s<MODE, false, TOPK, false>(mi, mu, a, grid, smem, st);
}

}  // namespace

// embedding_dim up to 128 (any value: the fp16 operand copies are zero-padded to K = 64 or 128)
bool mfb_tc_supported(const mfb_model *m, int k) {
  const int D = m->desc.dim;
  return D >= 1 && D <= 128 && k <= MFB_MAX_TOPK && m->items.rows >= 8 * TC_M && m->tune_tc != 0 &&
         ((uintptr_t)m->items.p & 31) == 0 && ((uintptr_t)m->users.p & 15) == 0;
}
static inline int tc_padded_dim(int D) { return D <= 64 ? 64 : 128; }

// Top-k for the listed users through the tensor-core path.  d_out_* as in mfb_topk.
// exact_topk: callback into the exact kernel for the (few) users that could not be certified.
int mfb_topk_tc(mfb_model *m, const int64_t *d_user_ids, int64_t n_users64, const int64_t *d_train_indptr,
                const int32_t *d_train_indices, int32_t k, int32_t *d_out_ids, float *d_out_scores, cudaStream_t st,
                int (*exact_topk)(mfb_model *, const int64_t *, int64_t, const int64_t *, const int32_t *, int32_t,
                                  int32_t *, float *, cudaStream_t, const int *, const int *),
                int *h_n_redo, uint64_t plan_key) {
  const int n_users = (int)n_users64;
  const int D = m->desc.dim, I = m->items.rows;
  const int n_users_pad = ((n_users + TC_N - 1) / TC_N) * TC_N;
  m->eval.n_users_pad = n_users_pad;
  const int i_tiles = (I + TC_M - 1) / TC_M;
  const int items_pad = i_tiles * TC_M;
  int sample_step = m->tune_tc_sample_step > 0 ? m->tune_tc_sample_step : 4;
  while ((i_tiles + sample_step - 1) / sample_step > 8 * TH_VPL) ++sample_step;   // groups = 4*n_sample <= 32*TH_VPL
  const int n_sample = (i_tiles + sample_step - 1) / sample_step;
  const int groups = n_sample * 4;
  // small user shards (user-sharded evaluation on many GPUs): split the item tiles of a user block over up to
  // TC_MAX_SPLIT CTAs so the grid still covers the SMs; every (user, split, column half) has its own candidate list
  int splits = 1;
  {
    const int user_ctas = n_users_pad / TC_N;
    while (splits < TC_MAX_SPLIT && user_ctas * (splits + 1) <= m->num_sms && n_sample >= 2 * (splits + 1)) ++splits;
    if (const char *e = getenv("MFB_TC_SPLIT")) if (*e) splits = atoi(e) < 1 ? 1 : (atoi(e) > TC_MAX_SPLIT ? TC_MAX_SPLIT : atoi(e));
  }
  EvalBuf &eb = m->eval;
  const int Dp = tc_padded_dim(D);
  MFB_CHECK(eb.ub.reserve((size_t)n_users_pad * Dp * sizeof(__half)));
  MFB_CHECK(eb.vb.reserve((size_t)items_pad * Dp * sizeof(__half)));
  MFB_CHECK(eb.unorm.reserve((size_t)n_users_pad * sizeof(float)));
  MFB_CHECK(eb.vnorm.reserve(((size_t)items_pad * 3 + i_tiles) * sizeof(float) + 16));
  MFB_CHECK(eb.gmax.reserve((size_t)(groups > 2 * TC_MAX_SPLIT * 24 ? groups : 2 * TC_MAX_SPLIT * 24) * n_users_pad * sizeof(int)));
  MFB_CHECK(eb.thr.reserve((size_t)n_users_pad * 2 * sizeof(float)));
  MFB_CHECK(eb.cand.reserve((size_t)n_users_pad * RS_MAXC * sizeof(int2)));
  MFB_CHECK(eb.cnt.reserve((size_t)n_users_pad * sizeof(int) * (2 * TC_MAX_SPLIT + 3) + 64));
  MFB_CHECK(eb.redo.reserve((size_t)n_users_pad * (sizeof(long long) + (size_t)k * (sizeof(int) + sizeof(float))) + 64));
  __half *ub = eb.ub.as<__half>(), *vb = eb.vb.as<__half>();
  float *unorm = eb.unorm.as<float>(), *vnorm = eb.vnorm.as<float>();
  float *vbias = vnorm + items_pad;   // item biases in position order
  float *vnorm_item = vbias + items_pad;   // scaled norms again, indexed by item id (for k_tc_rescore)
  float *tile_nmax = vnorm_item + items_pad;   // per tile: largest scaled norm
  float *thr = eb.thr.as<float>(), *eps = thr + n_users_pad;
  int *cand_cnt = eb.cnt.as<int>();
  int *redo_flag = cand_cnt + 2 * TC_MAX_SPLIT * (size_t)n_users_pad;
  int *redo_pos = redo_flag + n_users_pad;
  int *surv_cnt = redo_pos + n_users_pad;
  int *redo_cnt = surv_cnt + n_users_pad;

  int tk = m->prof.begin(PK_TOPK, st, 8);
  // Error model of the fp16 GEMM score against the exact fp32 score.  fp16 round-to-nearest: |x - fp16(x)| <= h|x| + s
  // with h = 2^-11 (11 significant bits) and s = 2^-25 (half a subnormal ulp; covers |x| < 6.1e-5), for |x| <= 65504.
  // Products of two fp16 numbers are exact in fp32, so
  //   |sum_d u_d v_d - sum_d fp16(u_d) fp16(v_d)| <= (2h + h^2) sum|u_d v_d| + s(1+h)(|u|_1 + |v|_1) + D s^2
  //                                               <= (2h + h^2) |u||v| + s(1+h) sqrt(D) (|u| + |v|) + D s^2
  // (Cauchy-Schwarz); fp32 accumulation in the tensor core and in the exact kernel adds < 2 * D * 2^-23 relative.
  //   c = 2^-10 * 1.075  (7% slack over 2h + h^2 + accumulation)      a = 2^-25 * sqrt(D) * 1.01
  // Stored per item: E_i = c|v_i| + a; per user: N_u = |u| + a/c.  Then E_i * N_u >= c|u||v_i| + a(|u| + |v_i|) + a^2/c,
  // which bounds the expression above (a^2/c > D s^2).  Rows with an element beyond the fp16 range raise `overflow`
  // and the whole call is served by the exact kernel.
  const float err_coeff = 1.0498e-3f;
  const float err_add = 2.98023224e-8f * sqrtf((float)D) * 1.01f;
  int *overflow = redo_cnt + 1;
  MFB_CUDA(cudaMemsetAsync(redo_cnt, 0, 2 * sizeof(int), st));
  k_tc_convert<<<(n_users_pad + 7) / 8, 256, 0, st>>>(m->users.p, (const long long *)d_user_ids, n_users, n_users_pad, D,
                                                     Dp, ub, unorm, 1.0f, err_add / err_coeff, 0, nullptr, nullptr, nullptr,
                                                     overflow);
  k_tc_convert<<<(items_pad + 7) / 8, 256, 0, st>>>(m->items.p, nullptr, I, items_pad, D, Dp, vb, vnorm, err_coeff, err_add,
                                                    i_tiles, m->items.bp, vbias, vnorm_item, overflow);
  k_tc_tile_maxnorm<<<(i_tiles + 7) / 8, 256, 0, st>>>(vnorm, i_tiles, tile_nmax);
  MFB_KERNEL_CHECK();
  // XK: bias and threshold through one extra K = 16 MMA step (operand images built here)
  const bool xk = m->tune_tc_xk != 0;
  const size_t ublocks = (size_t)n_users_pad / 128;
  uint8_t *xk_yimg = nullptr, *xk_ximg_max = nullptr, *xk_ximg_col = nullptr;
  float *xk_tprime = nullptr, *xk_rx = nullptr, *xk_rx_max = nullptr, *xk_bmax = nullptr;
  if (xk) {
    MFB_CHECK(eb.xk.reserve(((size_t)i_tiles + 2 * ublocks) * XK_BYTES + (3 * (size_t)n_users_pad + 4) * sizeof(float)));
    xk_yimg = eb.xk.as<uint8_t>();
    xk_ximg_max = xk_yimg + (size_t)i_tiles * XK_BYTES;
    xk_ximg_col = xk_ximg_max + ublocks * XK_BYTES;
    xk_tprime = reinterpret_cast<float *>(xk_ximg_col + ublocks * XK_BYTES);
    xk_rx = xk_tprime + n_users_pad;
    xk_rx_max = xk_rx + n_users_pad;
    xk_bmax = xk_rx_max + n_users_pad;
    MFB_CUDA(cudaMemsetAsync(xk_bmax, 0, sizeof(float), st));
    k_tc_xk_items<<<i_tiles, TC_M, 0, st>>>(vbias, i_tiles, xk_yimg, xk_bmax, overflow);
    k_tc_xk_users<<<(n_users_pad + 127) / 128, 128, 0, st>>>(n_users, n_users_pad, 0, nullptr, unorm, tile_nmax, i_tiles,
                                                            xk_bmax, xk_ximg_max, nullptr, xk_rx_max, overflow);
    MFB_KERNEL_CHECK();
  }

  CUtensorMap map_items, map_users;
  MFB_CHECK(make_tmap(&map_items, vb, items_pad, Dp, TC_M));
  MFB_CHECK(make_tmap(&map_users, ub, n_users_pad, Dp, TC_N));

  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.num_items = I;
  a.n_users = n_users;
  if (xk) {
    a.ximg = xk_ximg_max;
    a.yimg = xk_yimg;
    a.rad_extra = xk_rx_max;
  }
  a.D = Dp;
  a.total_tiles = i_tiles;
  a.item_bias = vbias;
  a.tile_nmax = tile_nmax;
  a.user_norm = unorm;
  a.n_users_pad = n_users_pad;
  if (const char *e = getenv("MFB_TC_DBG")) a.dbg = atoi(e);
#ifdef MFB_TC_TIMING
  static long long *d_timing = nullptr;
  if (!d_timing) MFB_CUDA(cudaMalloc(&d_timing, 48 * sizeof(long long)));
  a.timing = d_timing;
  auto dump_timing = [&](const char *what) {
    long long h[48];
    cudaStreamSynchronize(st);
    cudaMemcpy(h, d_timing, sizeof(h), cudaMemcpyDeviceToHost);
    for (int c = 0; c < 2; ++c) {
      const long long *t = h + 24 * c;
      const double nt = t[5] > 0 ? (double)t[5] : 1.0;
      fprintf(stderr, "[tc timing %s cta %d] tiles %lld | per tile: producer wait-empty %.0f | mma: wait-tempty %.0f wait-full %.0f issue %.0f total %.0f"
              " (%.3f GHz)"
              " | epi warp2: wait-tfull %.0f ld-wait %.0f proc(incl ld) %.0f prestore %.0f total %.0f | epi warp17: wait-tfull %.0f ld-wait %.0f proc %.0f prestore %.0f total %.0f\n",
              what, c, t[5], t[0] / nt, t[1] / nt, t[2] / nt, t[3] / nt, t[4] / nt, t[16] > 0 ? (double)t[4] / (double)t[16] : 0.0, t[6] / nt, t[7] / nt, t[8] / nt, t[9] / nt, t[10] / nt,
              t[11] / nt, t[12] / nt, t[13] / nt, t[14] / nt, t[15] / nt);
    }
  };
#endif
  // train mask for the COLLECT epilogue (dense bitmap) + per-user dirty sampled groups for the threshold kernel
  const int ncta = n_users_pad / TC_N;
  const uint32_t magic = ((uint64_t)items_pad * (uint64_t)i_tiles < (1ull << 32) && i_tiles > 1)
                             ? (uint32_t)(((1ull << 32) + (uint64_t)i_tiles - 1) / (uint64_t)i_tiles)
                             : 0u;
  int masked_in_gemm = 0;
  const size_t bitmap_bytes = (size_t)ncta * i_tiles * 4096;
  uint32_t *dirty = nullptr;
  if (d_train_indptr != nullptr && bitmap_bytes <= ((size_t)8 << 30) && groups <= 256 && k <= 32) {
    // The bitmap depends only on (user list, train CSR, item layout), not on the model: a caller that evaluates the
    // same interactions again (model.test() ranks them three times, implicit.py:428-460; every epoch of a validation
    // loop) names them with a non-zero plan_key and the bitmap of the previous call is reused.
    MFB_CHECK(eb.mpairs.reserve(bitmap_bytes));
    MFB_CHECK(eb.mptr.reserve((size_t)n_users_pad * 8 * sizeof(uint32_t)));
    dirty = eb.mptr.as<uint32_t>();
    const bool reuse = plan_key != 0 && eb.mask_key == plan_key && eb.mask_users == n_users && eb.mask_tiles == i_tiles &&
                       eb.mask_step == sample_step && eb.mask_pad == n_users_pad && eb.mask_bits_ptr == eb.mpairs.ptr &&
                       eb.mask_dirty_ptr == eb.mptr.ptr;   // (a grown buffer has lost the images)
    if (!reuse) {
      const size_t mb_smem = (size_t)(i_tiles < MB_TILES ? i_tiles : MB_TILES) * 512;
      MFB_CUDA(cudaFuncSetAttribute(k_tc_mask_bitmap, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)mb_smem));
      dim3 mgrid(ncta, 8);
      k_tc_mask_bitmap<<<mgrid, MB_THREADS, mb_smem, st>>>((const long long *)d_user_ids,
                                                           (const long long *)d_train_indptr, d_train_indices, n_users,
                                                           i_tiles, magic, sample_step, eb.mpairs.as<uint32_t>(), dirty);
      MFB_KERNEL_CHECK();
      eb.mask_key = plan_key;
      eb.mask_users = n_users;
      eb.mask_tiles = i_tiles;
      eb.mask_step = sample_step;
      eb.mask_pad = n_users_pad;
      eb.mask_bits_ptr = eb.mpairs.ptr;
      eb.mask_dirty_ptr = eb.mptr.ptr;
    }
    a.mask_bits = eb.mpairs.as<uint2>();
    masked_in_gemm = 1;
  }
  // sample pass: group maxima of every sample_step-th item tile (no train mask: the small-k threshold kernel drops
  // the groups that contain a train item of the user; the general one counts k + ntrain maxima instead)
  const bool small_thr = k <= 32 && groups <= 256 && (d_train_indptr == nullptr || masked_in_gemm);
  a.tile_begin = 0;
  a.tile_step = sample_step;
  a.n_tiles = n_sample;
  a.gmax = eb.gmax.as<int>();
  // k <= 24 (and at most 256 sampled groups): the MAX epilogue keeps each thread's 24 largest clean group maxima and a
  // small kernel merges a user's 2 * splits lists; otherwise the maxima go through gmax and a selection kernel
  constexpr int THRK = 24;
  const bool fused_thr = small_thr && k <= THRK && m->tune_tc_fused_thr != 0;
  // User-block ranges.  All CTAs of a pass do the same work, so U user blocks on S SMs take ceil(U / S) waves for U / S
  // waves of work (cfg4: 541 blocks on 148 SMs = 4 waves for 3.66).  With the fused threshold path the whole waves
  // run as they are and the blocks of the partial last wave have their item tiles split over 2-4 CTAs each, so that it
  // takes ceil(tail * splits / S) / splits of a wave (cfg4: 97 blocks x 3 = 291 CTAs = 2/3 of a wave).  A range has its
  // own number of candidate sub-lists per user (2 * splits, RS_MAXC / (2 * splits) records each).  Measured at cfg4:
  // 1.556 vs 1.567 ms per pass -- the board is at its power cap, so the SMs of a partial wave run at a higher clock and the
  // idle ones cost little -- hence off by default (MFB_TC_TAIL_SPLIT=1).
  struct Range { int block0, nblocks, splits; };
  Range rg[2] = {{0, n_users_pad / TC_N, splits}, {0, 0, 1}};
  int nrg = 1;
  {
    const int user_ctas = n_users_pad / TC_N, sms = m->num_sms;
    const int tail = user_ctas % sms;
    const char *forced = getenv("MFB_TC_SPLIT");
    if (fused_thr && m->tune_tc_tail_split != 0 && !(forced && *forced) && user_ctas > sms && tail != 0) {
      int best = 1;
      double best_cost = 1.0;
      for (int s2 = 2; s2 <= TC_MAX_SPLIT && n_sample >= 2 * s2; ++s2) {
        const double cost = (double)((tail * s2 + sms - 1) / sms) / s2;
        if (cost < best_cost - 1e-9) {
          best_cost = cost;
          best = s2;
        }
      }
      if (best > 1) {
        rg[0] = {0, user_ctas - tail, 1};
        rg[1] = {user_ctas - tail, tail, best};
        nrg = 2;
      }
    }
  }
  eb.n_ranges = nrg;
  for (int r = 0; r < nrg; ++r) {
    eb.range_first[r] = rg[r].block0 * TC_N;
    const int end = (rg[r].block0 + rg[r].nblocks) * TC_N;
    eb.range_users[r] = (end < n_users ? end : n_users) - eb.range_first[r];
    eb.range_nsub[r] = 2 * rg[r].splits;
  }
  if (fused_thr) {
    a.toplists = eb.gmax.as<int>();   // (same buffer: reserved for max(groups, 2 * TC_MAX_SPLIT * THRK) rows)
    a.dirty_groups = dirty;
    for (int r = 0; r < nrg; ++r) {
      a.block_base = rg[r].block0;
      MFB_CHECK((launch_gemm<MODE_MAX, THRK>(map_items, map_users, a, rg[r].nblocks, st, rg[r].splits)));
    }
  } else {
    MFB_CHECK(launch_gemm<MODE_MAX>(map_items, map_users, a, rg[0].nblocks, st, splits));
  }
#ifdef MFB_TC_TIMING
  dump_timing("MAX");
#endif
  if (fused_thr) {
    for (int r = 0; r < nrg; ++r) {
      const int first = eb.range_first[r], nu_r = eb.range_users[r];
      if (nu_r <= 0) continue;
      k_tc_threshold_merge<THRK><<<(nu_r + 127) / 128, 128, 0, st>>>(eb.gmax.as<int>() + first, eb.range_nsub[r], nu_r,
                                                                     n_users_pad, k, thr + first);
    }
  } else if (small_thr) {
    int *gm = eb.gmax.as<int>();
    // (four threads per user: measured SLOWER at 17 312 and 34 624 users -- the two merge rounds cost more than the
    // latency they hide -- so it is off unless MFB_TC_THR_TPU4 names a user count below which to use it)
    int tpu4_below = 0;
    if (const char *e = getenv("MFB_TC_THR_TPU4")) tpu4_below = atoi(e);
#define MFB_THRESHOLD(K)                                                                                             \
  do {                                                                                                               \
    if (n_users < tpu4_below)                                                                                        \
      k_tc_threshold_small<K, 4><<<(4 * n_users + 127) / 128, 128, 0, st>>>(gm, groups, n_users, n_users_pad, k, thr, \
                                                                            dirty);                                  \
    else                                                                                                             \
      k_tc_threshold_small<K, 1><<<(n_users + 127) / 128, 128, 0, st>>>(gm, groups, n_users, n_users_pad, k, thr,     \
                                                                        dirty);                                      \
  } while (0)
    if (k <= 8) MFB_THRESHOLD(8);
    else if (k <= 16) MFB_THRESHOLD(16);
    else if (k <= 24) MFB_THRESHOLD(24);
    else MFB_THRESHOLD(32);
#undef MFB_THRESHOLD
  } else {
    k_tc_threshold<<<(n_users + 3) / 4, 128, 0, st>>>(eb.gmax.as<int>(), groups, n_users, n_users_pad, k,
                                                      (const long long *)d_user_ids,
                                                      (const long long *)d_train_indptr, unorm, nullptr, thr, eps,
                                                      0);
  }
  MFB_KERNEL_CHECK();
  // collect pass: all tiles
  if (xk) {
    k_tc_xk_users<<<(n_users_pad + 127) / 128, 128, 0, st>>>(n_users, n_users_pad, 1, thr, unorm, tile_nmax, i_tiles, xk_bmax,
                                                            xk_ximg_col, xk_tprime, xk_rx, overflow);
    MFB_KERNEL_CHECK();
    a.ximg = xk_ximg_col;
    a.rad_extra = xk_rx;
  }
  a.tile_begin = 0;
  a.tile_step = 1;
  a.n_tiles = i_tiles;
  a.thr = thr;
  a.cand = eb.cand.as<int2>();
  a.cand_cnt = cand_cnt;
  for (int r = 0; r < nrg; ++r) {
    a.block_base = rg[r].block0;
    a.cap2 = RS_MAXC / eb.range_nsub[r];
    MFB_CHECK(launch_gemm<MODE_COLLECT>(map_items, map_users, a, rg[r].nblocks, st, rg[r].splits));
  }
#ifdef MFB_TC_TIMING
  dump_timing("COLLECT");
#endif
  // exact re-score + mask + top-k
  long long *redo_users = eb.redo.as<long long>();
  const size_t rs_smem = (size_t)RS_WARPS * (D + 3 * RS_MAXC) * sizeof(float);
#define MFB_RESCORE(MAXSUB)                                                                                          \
  do {                                                                                                               \
    static size_t rs_set = 0;                                                                                        \
    if (rs_smem > rs_set) {                                                                                          \
      MFB_CUDA(cudaFuncSetAttribute(k_tc_rescore<MAXSUB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)rs_smem)); \
      rs_set = rs_smem;                                                                                              \
    }                                                                                                                \
    k_tc_rescore<MAXSUB><<<(nu_r + RS_WARPS - 1) / RS_WARPS, RS_WARPS * 32, rs_smem, st>>>(                          \
        (const long long *)d_user_ids + first, nu_r, m->users, m->items, D, eb.cand.as<int2>() + (size_t)first * RS_MAXC, \
        cand_cnt + (size_t)first * TC_CNT_STRIDE, RS_MAXC / nsub_r, nsub_r, thr + first, unorm + first, vnorm_item,   \
        (const long long *)d_train_indptr, d_train_indices, k, d_out_ids + (size_t)first * k,                         \
        d_out_scores ? d_out_scores + (size_t)first * k : nullptr, redo_flag + first, surv_cnt + first,               \
        masked_in_gemm ? 0 : 1, redo_users, redo_pos, redo_cnt, xk ? xk_tprime + first : nullptr,                     \
        xk ? xk_rx + first : nullptr, m->tune_tc_tile_radius ? tile_nmax : nullptr, i_tiles, magic, first);           \
  } while (0)
  for (int r = 0; r < nrg; ++r) {
    const int first = eb.range_first[r], nu_r = eb.range_users[r], nsub_r = eb.range_nsub[r];
    if (nu_r <= 0) continue;
    if (nsub_r == 2) MFB_RESCORE(2);
    else if (nsub_r <= 4) MFB_RESCORE(4);
    else MFB_RESCORE(2 * TC_MAX_SPLIT);
  }
#undef MFB_RESCORE
  MFB_KERNEL_CHECK();
  m->prof.end(tk, st);
  // The flagged users are re-done by the exact kernel straight into their rows of the result.  Their number stays on
  // the device (the launch covers the worst case with a grid of striding blocks that leave at once when there is
  // nothing to do), so the whole pass is enqueued without a host synchronisation; mfb_topk_last_redo reads it on demand.
  MFB_CHECK(exact_topk(m, (const int64_t *)redo_users, n_users, d_train_indptr, d_train_indices, k, d_out_ids, d_out_scores,
                       st, redo_cnt, redo_pos));
  eb.redo_cnt_dev = redo_cnt;
  eb.redo_stream = st;
  if (h_n_redo) *h_n_redo = -1;   // pending
  return MFB_OK;
}

// debug: candidate-list statistics of the last tensor-core top-k call: {users, sum, max, over_cap, re-scored}
int mfb_tc_stats(mfb_model *m, int n_users, long long *h_out, cudaStream_t st) {
  const int n_users_pad = m->eval.n_users_pad;
  const EvalBuf &eb = m->eval;
  std::vector<int> cnt((size_t)n_users * TC_CNT_STRIDE);
  MFB_CUDA(cudaMemcpyAsync(cnt.data(), m->eval.cnt.ptr, cnt.size() * sizeof(int), cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  long long sum = 0, mx = 0, over = 0;
  for (int r = 0; r < eb.n_ranges; ++r) {
    const int nsub = eb.range_nsub[r];
    for (int u = eb.range_first[r]; u < eb.range_first[r] + eb.range_users[r] && u < n_users; ++u) {
      long long tot = 0;
      bool ov = false;
      for (int j = 0; j < nsub; ++j) {
        const int c = cnt[(size_t)u * TC_CNT_STRIDE + j];
        tot += c;
        ov |= c > RS_MAXC / nsub;
      }
      sum += tot;
      if (tot > mx) mx = tot;
      if (ov) ++over;
    }
  }
  cnt.resize((size_t)n_users);
  h_out[0] = n_users;
  h_out[1] = sum;
  h_out[2] = mx;
  h_out[3] = over;
  // listed items that survived the bound filter and were re-scored exactly
  MFB_CUDA(cudaMemcpyAsync(cnt.data(), m->eval.cnt.as<int>() + (2 * TC_MAX_SPLIT + 2) * (size_t)n_users_pad, (size_t)n_users * sizeof(int),
                           cudaMemcpyDeviceToHost, st));
  MFB_CUDA(cudaStreamSynchronize(st));
  long long surv = 0;
  for (int u = 0; u < n_users; ++u) surv += cnt[(size_t)u];
  h_out[4] = surv;
  return MFB_OK;
}

// debug / test hook: dump the raw tensor-core scores (item-major) of the listed users
int mfb_tc_dump_scores(mfb_model *m, const int64_t *d_user_ids, int n_users, float *d_out, cudaStream_t st) {
  const int D = m->desc.dim, I = m->items.rows;
  const int n_users_pad = ((n_users + TC_N - 1) / TC_N) * TC_N;
  const int i_tiles = (I + TC_M - 1) / TC_M;
  const int items_pad = i_tiles * TC_M;
  EvalBuf &eb = m->eval;
  const int Dp = tc_padded_dim(D);
  MFB_CHECK(eb.ub.reserve((size_t)n_users_pad * Dp * sizeof(__half)));
  MFB_CHECK(eb.vb.reserve((size_t)items_pad * Dp * sizeof(__half)));
  MFB_CHECK(eb.unorm.reserve((size_t)n_users_pad * sizeof(float)));
  MFB_CHECK(eb.vnorm.reserve((size_t)items_pad * 3 * sizeof(float) + 16));
  MFB_CHECK(eb.cnt.reserve(64));
  __half *ub = eb.ub.as<__half>(), *vb = eb.vb.as<__half>();
  k_tc_convert<<<(n_users_pad + 7) / 8, 256, 0, st>>>(m->users.p, (const long long *)d_user_ids, n_users, n_users_pad, D,
                                                     Dp, ub, eb.unorm.as<float>(), 1.0f, 0.f, 0, nullptr, nullptr, nullptr,
                                                     eb.cnt.as<int>());
  k_tc_convert<<<(items_pad + 7) / 8, 256, 0, st>>>(m->items.p, nullptr, I, items_pad, D, Dp, vb, eb.vnorm.as<float>(), 1.0f,
                                                    0.f, i_tiles, m->items.bp, eb.vnorm.as<float>() + items_pad, nullptr,
                                                    eb.cnt.as<int>());
  MFB_KERNEL_CHECK();
  CUtensorMap map_items, map_users;
  MFB_CHECK(make_tmap(&map_items, vb, items_pad, Dp, TC_M));
  MFB_CHECK(make_tmap(&map_users, ub, n_users_pad, Dp, TC_N));
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.num_items = I;
  a.n_users = n_users;
  a.D = Dp;
  a.total_tiles = i_tiles;
  a.item_bias = eb.vnorm.as<float>() + items_pad;
  a.n_users_pad = n_users_pad;
  if (m->tune_tc_xk != 0) {
    // the same extra-K-step operands as the top-k passes use (threshold pieces zero): the dump then shows the score
    // arithmetic of the production kernels, bias pieces included
    const size_t ublocks = (size_t)n_users_pad / 128;
    MFB_CHECK(eb.xk.reserve(((size_t)i_tiles + ublocks) * XK_BYTES + ((size_t)n_users_pad + 4) * sizeof(float)));
    uint8_t *yimg = eb.xk.as<uint8_t>();
    uint8_t *ximg = yimg + (size_t)i_tiles * XK_BYTES;
    float *rx = reinterpret_cast<float *>(ximg + ublocks * XK_BYTES);
    float *bmax = rx + n_users_pad;
    MFB_CUDA(cudaMemsetAsync(bmax, 0, sizeof(float), st));
    k_tc_xk_items<<<i_tiles, TC_M, 0, st>>>(eb.vnorm.as<float>() + items_pad, i_tiles, yimg, bmax, eb.cnt.as<int>());
    k_tc_xk_users<<<(n_users_pad + 127) / 128, 128, 0, st>>>(n_users, n_users_pad, 0, nullptr, eb.unorm.as<float>(),
                                                            nullptr, i_tiles, bmax, ximg, nullptr, rx, eb.cnt.as<int>());
    MFB_KERNEL_CHECK();
    a.ximg = ximg;
    a.yimg = yimg;
  }
  a.tile_begin = 0;
  a.tile_step = 1;
  a.n_tiles = i_tiles;
  a.dump = d_out;   // [I][n_users_pad]
  return launch_gemm<MODE_DUMP>(map_items, map_users, a, n_users_pad / TC_N, st);
}
