// Shared definitions for the MoE prefill kernels (layouts, work descriptors).
#pragma once
#include <cstdint>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include <atomic>

namespace kb2 {

// cudaFuncAttributeMaxDynamicSharedMemorySize is a PER-DEVICE attribute: one process may drive several GPUs (the
// reference's EP layout, python/krasis/model.py:1496-1523), so the opt-in is remembered per device ordinal, not per
// process.  Thread-safe (atomic flags; setting the attribute twice is harmless).
struct PerDeviceOnce {
  std::atomic<bool> done[64];
  PerDeviceOnce() { for (auto& d : done) d.store(false); }
  // returns the current device ordinal if its opt-in is still pending, else -1
  int pending() {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 0;   // unknown device: always (re)configure
    return done[dev].load(std::memory_order_acquire) ? -1 : dev;
  }
  void mark(int dev) { if (dev >= 0 && dev < 64) done[dev].store(true, std::memory_order_release); }
};

// ---------------------------------------------------------------------------------------------
// B200 expert-weight layout ("KB2 tiles").  The reference's quantiser emits, per expert and matrix,
// packed[N][K/8] u32 (nibble i of word w = column 8w+i, value q+8) and scales[N][K/128] bf16
// (src/weights/marlin.rs:145-207); its GPU form is a Marlin permutation of those values
// (marlin.rs:323-491) designed for mma.sync.  For tcgen05 we keep the same VALUES and re-tile:
//
//   packed tile (nt, kb) = 128 rows x 64 K-columns = 4096 B at ((nt * K/64) + kb) * 4096
//       [half h in {0,1}] [row r in 0..127] [16 B]   16 B = 4 words, word j covers columns
//       kb*64 + h*32 + j*8 + {0,2,4,6,1,3,5,7}  (nibble p holds column offset kNibOrder[p]) so that
//       (w >> 4i) & 0x000F000F yields the BF16 pair (col 2i, col 2i+1) directly.
//   scale  tile (nt, g)  = 128 x bf16 = 256 B at ((nt * K/128) + g) * 256,  [row r]
//
// One tile is one contiguous blob => a single 1-D bulk-TMA copy per (tile, k-block), fully coalesced,
// and bank-conflict-free 16 B reads per thread-row in the dequant warps.
// ---------------------------------------------------------------------------------------------
constexpr int kTileRows = 128;
constexpr int kBlockK = 64;
constexpr int kGroup = 128;
constexpr int kInt4TileBytes = kTileRows * kBlockK / 2;   // 4096
constexpr int kInt8TileBytes = kTileRows * kBlockK;       // 8192
constexpr int kScaleTileBytes = kTileRows * 2;            // 256

enum WeightFormat : int {
  kFmtInt4G128 = 0,   // Krasis symmetric INT4, group 128
  kFmtInt8G128 = 1,   // Krasis symmetric INT8, group 128
  kFmtQ8_0 = 2,       // GGUF Q8_0 blocks (fp16 d + 32 x i8), re-tiled losslessly
  kFmtQ4_K = 3,       // GGUF Q4_K super-blocks (fp16 d, dmin, 6-bit scales/mins, 4-bit quants), re-tiled losslessly
  kFmtAffine8 = 4,    // kernel-side form of the GGUF types below: int8 codes + per-16-element (a, b) f32, w = fma(a, code, -b)
  // API-level ids of the GGUF types decoded into kFmtAffine8 tiles at load time (lossless: same values, same single f32 rounding)
  kFmtQ6_K = 4,       // 210 B / 256: w = (d*sc)*(q-32), sc per 16 elements            (src/gguf.rs:813-866, gguf_kernels.rs:594-635)
  kFmtQ5_K = 5,       // 176 B / 256: w = (d*sc)*q - dmin*mn, 5-bit q                   (src/gguf.rs:740-811)
  kFmtQ5_0 = 6,       // 22 B / 32:   w = d*(q-16)                                      (src/gguf.rs:599-633)
  kFmtQ4_0 = 7,       // 18 B / 32:   w = d*(q-8)                                       (src/gguf.rs:635-664)
};
__host__ __device__ constexpr int kernel_format(int fmt) { return fmt >= kFmtAffine8 ? kFmtAffine8 : fmt; }

// GGUF tile blobs are self-contained (block scales travel with the quants):
//   Q8_0 (128 rows x 64 K): [quarter 0..3][row][16 B] int8 (8192 B) | [row][2] fp16 d of the two 32-element blocks (512 B)
//   Q4_K (128 rows x 64 K = one 64-element chunk j of the 256-element super-block, src/gguf.rs:681-738):
//        [half 0..1][row][16 B] qs bytes (byte l: low nibble = element l, high nibble = element 32+l) (4096 B)
//        | [row][8 B] = fp16 d, fp16 dmin, u8 sc_lo, mn_lo, sc_hi, mn_hi (get_scale_min_k4 already applied) (1024 B)
constexpr int kQ8_0TileBytes = kTileRows * kBlockK + kTileRows * 4;   // 8704
constexpr int kQ4KTileBytes = kTileRows * kBlockK / 2 + kTileRows * 8; // 5120
//   Affine8 (128 rows x 64 K): [quarter 0..3][row][16 B] int8 codes (8192 B) | [row][4 x (a f32, b f32)] one pair per 16 elements (4096 B)
constexpr int kAffine8TileBytes = kTileRows * kBlockK + kTileRows * 32;  // 12288

// One unit of grouped-GEMM work along the token axis: a run of <= kMaxChunkTokens sorted slots of one expert.
struct ChunkDesc {
  int expert;      // local expert index
  int slot_begin;  // first sorted slot
  int n_tok;       // real token slots in this chunk (1..kMaxChunkTokens)
  int pad_;
};

constexpr int kMaxChunkTokens = 192;   // UMMA N per chunk; 2 accumulators x 192 + 2 A stages x 64 = 512 TMEM columns

struct GemmParams {
  const uint8_t* wq;          // packed tiles, all local experts
  const uint8_t* ws;          // scale tiles
  long long wq_expert_stride; // bytes
  long long ws_expert_stride; // bytes
  int n_kblocks;              // K / 64
  int items_per_chunk;        // GEMM1: I/128 (gate tile + up tile)   GEMM2: H/256 (two consecutive tiles)
  int tile1_offset;           // GEMM1: I/128   GEMM2: 1
  int tile0_mul;              // GEMM1: 1       GEMM2: 2          tile0 = rt * tile0_mul, tile1 = tile0 + tile1_offset
  const ChunkDesc* chunks;
  const int* n_chunks;        // device scalar
  __nv_bfloat16* out;         // GEMM1: act[slot][I]   GEMM2: c3[slot][H]
  long long out_ld;
  const float* slot_weight;   // GEMM2: routing weight per sorted slot
  const int* gather_rows;     // GEMM1, gather mode: token row of every sorted slot (the B operand is fetched from the caller's
                              // activation matrix with TMA gather4 instead of a pre-sorted copy); nullptr = B rows are contiguous
  long long* trace;           // optional (tuning only, KB2_GEMM_TRACE=<device pointer>): clock64 stamps of CTA 0, [item < 16][16]
};

// combine_kernel output redirection for the fused reduce-scatter of expert parallelism (n_ranks == 0: plain local output)
constexpr int kMaxPeers = 8;
struct CombineScatter {
  __nv_bfloat16* peer_out[kMaxPeers] = {};   // receive buffer of every rank ([rows_per_rank][n_ranks][H] bf16), peer-mapped
  int rows_per_rank = 0, src_rank = 0, n_ranks = 0;
};

}  // namespace kb2
