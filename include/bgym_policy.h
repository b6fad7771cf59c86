/* bgym_policy.h — C-ABI of the fused rollout-policy forward (libbgym_policy.so).
 *
 * Policy side of on-device PPO rollout collection (BASELINE configs[4]); NOT part of the env step
 * path and built as its own library so that the env library stays free of tensor-core code.
 *
 * The network is the reference's extractor + SB3 heads (train_balatro_agent.py:48-81, :341; widths as the data has
 * them, see balatro_gym_b200/rollout.py::make_policy):
 *     hand_net  416 -> 256 -> 128      joker_net 10 -> 128 -> 64      game_state_net 21 -> 64 -> 32      (ReLU)
 *     combined_net 224 -> 512 -> 512 (ReLU)      pi 512 -> 256 -> 256 -> 60 (tanh)      vf 512 -> 256 -> 256 -> 1 (tanh)
 * This entry point runs all FOURTEEN layers in ONE kernel on the 5th-generation tensor cores (tcgen05.mma, bf16 inputs,
 * fp32 accumulators in tensor memory), straight from the observation records: a CTA owns a tile of 128 envs, builds the
 * 8-hot hand block / joker ids / scaled game scalars of the tile in shared memory (128-byte-swizzled K-major, the layout
 * the MMA reads), keeps the activations there from the first layer to the last, and streams the 1.9 MB of weights through
 * a three-stage ring of 1-D bulk copies; between layers the accumulators come back through tcgen05.ld for bias +
 * activation and go straight into the next layer's operand buffer.  Nothing but the 176-byte observation record and the
 * 60 logits + value per env touches HBM (the library-GEMM path moves ~12 KB of activations per env).
 */
#ifndef BGYM_POLICY_H
#define BGYM_POLICY_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BGYM_POLICY_IN_DIM   448   /* relu([hand 256 | joker 128 | game 64]): the output of bgym_policy_first_layer */
#define BGYM_POLICY_LOGITS   60
#define BGYM_POLICY_LAYERS   14
#define BGYM_POLICY_MAX_STEPS 80

/* One weight tile of the kernel's program: rows [n0, n0 + n) x input columns [64 kb, 64 kb + 64) of layer `layer`,
 * stored at byte `offset` of the packed weight blob as n rows of 128 bytes, 128-byte swizzled (16-byte chunk c of row r
 * at chunk position c ^ (r & 7)); rows / columns beyond the layer's real shape are zero. */
typedef struct BgymPolicyStep {
  int32_t offset;     /* byte offset in the packed blob          */
  int32_t bytes;      /* n * 128                                 */
  int32_t layer;      /* 0 hand_net.0, 1 joker_net.0, 2 game_state_net.0 (input columns shifted by 16: see the .cu),
                         3 hand_net.2, 4 joker_net.2, 5 game_state_net.2, 6 combined_net.0, 7 combined_net.2,
                         8 pi.0, 9 vf.0, 10 pi.2, 11 vf.2, 12 pi.4, 13 vf.4                                      */
  int32_t n0;         /* first output row of the tile            */
  int32_t n;          /* rows in the tile (MMA N): 16..256       */
  int32_t kb;         /* 64-column block of the layer's input    */
  int32_t a_kb;       /* where that block sits in the activation buffer */
  int32_t col;        /* first accumulator column (tensor memory) */
  int32_t first;      /* 1 = first K block of this output tile (overwrite the accumulator) */
  int32_t last;       /* 1 = last tile before an epilogue        */
  int32_t group;      /* epilogue group 0..6                     */
  int32_t _pad;
} BgymPolicyStep;

/* the program (same for every call): fills steps[BGYM_POLICY_MAX_STEPS], returns the number of steps; *weight_bytes /
 * *bias_floats (either may be NULL) receive the sizes of the packed blobs.  Bias blob: per epilogue group, one float per
 * accumulator column — group g starts at float offset 512 g. */
int bgym_policy_program(BgymPolicyStep* steps, int64_t* weight_bytes, int64_t* bias_floats);

/* obs: n BgymObs records (176 B each, include/bgym.h; only fields a step keeps current in the records are read: hand,
 * joker_ids and the game scalars — no bgym_sync_obs needed); weights / bias: the packed blobs (device); logits: n x 60 fp32;
 * value: n fp32.  Asynchronous on `stream`.  Returns 0, a cudaError_t (> 0) or -1 for bad arguments. */
int bgym_policy_forward(const void* obs, const void* weights, const float* bias, float* logits, float* value,
                        int64_t n, void* stream);
const char* bgym_policy_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
