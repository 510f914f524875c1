// otter_b200 — definitions shared by the fused attention kernels (csrc/otb_attn.cu, csrc/otb_xattn_fused.cu):
// problem description, key-tile / mask-row helpers, SW128 staging store.
#pragma once
#include "otb_common.cuh"

namespace otb {

constexpr int kAttnThreads = 128;
constexpr int kTileBytes = 128 * 128;  // 128 rows x 64 bf16 (128 B) = 16 KB

struct AttnParams {
  int P, H, Sq, Sk1, Sk2;
  int q_col0, k1_col0, v1_col0, k2_col0, v2_col0;
  float scale, scale_log2;
  // forward outputs
  bf16* out; long long ldo; int o_col0;
  float* lse;  // [P][H][Sq]
  // mask (mode B) — text_time == nullptr disables masking; mask_ge: keys of every media slot <= text_time
  // (only_attend_immediate_media=False, modeling_otter.py:317 mask_op = torch.ge); causal: keys j <= query row
  const int* text_time; int n_per_media; int T_img; int mask_ge; int causal;
  // backward
  const bf16* o; const bf16* dout; long long ld_do; int do_col0;
  bf16* dq; long long ld_dq; int dq_col0;
  bf16* dkv1; long long ld_dkv1; int dk1_col0, dv1_col0;
  bf16* dkv2; long long ld_dkv2; int dk2_col0, dv2_col0;
  float* dq_ws;  // [P*Sq][H*64] fp32, only touched when more than one key tile
};

struct KeyTile {
  int src, row0, valid, key_base;
};
__device__ __forceinline__ KeyTile key_tile(const AttnParams& p, int prob, int j, int nt1) {
  KeyTile t;
  if (j < nt1) {
    t.src = 0; t.row0 = prob * p.Sk1 + j * 128; t.valid = min(128, p.Sk1 - j * 128); t.key_base = j * 128;
  } else {
    const int jj = j - nt1;
    t.src = 1; t.row0 = prob * p.Sk2 + jj * 128; t.valid = min(128, p.Sk2 - jj * 128); t.key_base = p.Sk1 + jj * 128;
  }
  return t;
}

// row class: 0 = zeroed row, 1 = normal, 2 = uniform (fully masked)
__device__ __forceinline__ int row_class(const AttnParams& p, int tt) {
  if (p.text_time == nullptr) return 1;
  // torch.ge: a row before the first <image> has every key masked -> uniform; it is NOT zeroed afterwards
  // (the zeroing at modeling_otter.py:326-330 is guarded by only_attend_immediate_media)
  if (p.mask_ge) return (tt == 0) ? 2 : 1;
  if (tt == 0) return 0;
  return (tt <= p.T_img) ? 1 : 2;
}
__device__ __forceinline__ bool key_allowed(const AttnParams& p, int tt, int key_idx) {
  if (p.text_time == nullptr) return true;
  return (key_idx / p.n_per_media + 1) == tt;
}

// Per-row key window inside one key tile, replacing a per-element mask test (and its integer division):
//   class 0 -> empty; class 2 (uniform) -> every valid key with weight 1; class 1 -> exp() over [lo, hi)
struct RowRange {
  int lo, hi;
};
__device__ __forceinline__ RowRange row_range(const AttnParams& p, int cls, int tt, const KeyTile& kt, int row) {
  RowRange r;
  r.lo = 0;
  r.hi = (cls == 0) ? 0 : kt.valid;
  if (cls == 1 && p.text_time != nullptr) {
    if (p.mask_ge) {                       // media slots 1 .. min(tt, T): one contiguous key prefix
      r.hi = max(0, min(kt.valid, min(tt, p.T_img) * p.n_per_media - kt.key_base));
    } else {
      const int s0 = (tt - 1) * p.n_per_media - kt.key_base;
      r.lo = max(0, s0);
      r.hi = max(r.lo, min(kt.valid, s0 + p.n_per_media));
    }
  }
  if (p.causal) r.hi = max(r.lo, min(r.hi, row + 1 - kt.key_base));   // self-attention: key index <= query index
  return r;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// write 8 consecutive bf16 (one 16 B unit) of row r, columns [c, c+8) into a [rows][64] SW128 tile chunk
__device__ __forceinline__ void st_sw128(uint8_t* chunk_base, int r, int c_in_chunk, uint4 v) {
  const int unit = (c_in_chunk >> 3) ^ (r & 7);
  *reinterpret_cast<uint4*>(chunk_base + r * 128 + unit * 16) = v;
}

}  // namespace otb
