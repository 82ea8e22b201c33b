/*
 * htsp_b200 -- C ABI of the B200-native (sm_100a) H-TSP lower-level path-solver hot path.
 *
 * The reference (Learning4Optimization-HUST/H-TSP) is pure Python/PyTorch and has no FFI:
 * its boundary for this path is two duck-typed Python interfaces, RLSolver.solve
 * (h_tsp.py:165-228) and PathSolver.forward (rl4cop/train_path_solver.py:224-311).  The
 * entry points below are what a ctypes binding on the reference side calls instead of the
 * eager ATen op sequences cited on each function (INTEGRATION.md shows the stub).
 *
 * Conventions: plain pointers and sizes only (no torch types); every pointer that is not
 * explicitly "host" is a DEVICE pointer; all work is enqueued on `stream` (a cudaStream_t
 * passed as void*) and nothing synchronises; no hidden allocation -- scratch memory is a
 * caller-provided workspace sized by the matching *_workspace_bytes query; the return
 * value is 0 on success and a negative htsp_status otherwise (htsp_last_error() has text).
 * Shapes: B subproblems, A = 1 or 8 augmentations, Bp = A*B instances, N nodes per
 * instance, G POMO starts (rows) per instance, H = 128, 8 heads x 16, FF 512.
 */
#ifndef HTSP_B200_H_
#define HTSP_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HTSP_ABI_VERSION 1
#define HTSP_EMBED_DIM 128
#define HTSP_N_HEADS 8
#define HTSP_FF_DIM 512

typedef enum {
  HTSP_OK = 0,
  HTSP_ERR_INVALID_ARG = -1,
  HTSP_ERR_WORKSPACE = -2,
  HTSP_ERR_CUDA = -3,
  HTSP_ERR_UNSUPPORTED = -4
} htsp_status;

/* arithmetic of the decode/encode kernels */
typedef enum {
  HTSP_MODE_FP32 = 0,   /* fp32 FMA on CUDA cores, reference op order (parity anchor)        */
  HTSP_MODE_TC_X3 = 1,  /* tensor cores, split-fp16 operands (hi+lo, 3 MMAs), fp32 accumulate */
  HTSP_MODE_TC_FAST = 2 /* as X3 with one single-fp16 operand in the glimpse (tcgen05 kernel: the
                           softmax probabilities; mma.sync kernel: QK^T); logits within 1e-3     */
} htsp_mode;

int htsp_abi_version(void);
const char* htsp_last_error(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
uint64_t htsp_launch_count(void);

/* Measurement hook: when both are non-NULL cudaEvent_t handles, the NEXT launch of the dominant
 * kernel (the rollout kernel inside htsp_decode) is bracketed by cudaEventRecord(start) /
 * cudaEventRecord(stop) on its stream; the pair is consumed by that launch.  bench.py uses
 * this for the live roofline number.                                                       */
int htsp_set_kernel_timer(void* start_event, void* stop_event);

/* ---- weights ------------------------------------------------------------------------
 * Replaces PathSolver.load_state_dict + the per-call nn.Linear / BatchNorm1d parameter reads
 * (rl4cop/models/models.py:12-60, 331-358).  `host_tensors` is a HOST array of HOST fp32
 * pointers in the canonical order of htsp_b200.weights.expected_keys(n_layers).  Eval-mode
 * BatchNorm (models.py:35,37, eps 1e-5) is folded to per-channel alpha/beta; Wq/Wk/Wv and
 * the decoder projections are concatenated; multi_head_combine is also stored transposed.
 * The packed blob (htsp_weights_blob_bytes) is written to device memory `dev_blob`.     */
size_t htsp_weights_blob_bytes(int n_layers);
int htsp_n_weight_tensors(int n_layers);
int htsp_pack_weights(const float* const* host_tensors, int n_tensors, int n_layers,
                      void* dev_blob, void* stream);

/* ---- K1: subproblem extraction ------------------------------------------------------
 * RLSolver.solve gather + min/max normalisation into [0.05,0.95]^2 (h_tsp.py:176-191) and
 * utils.augment_xy_data_by_8_fold (rl4cop/utils/utils.py:37-56, block-major).
 * data [B,Ntot,2] f32, fragment [B,N] i64 -> x_new [B,N,2], scale [B] (= s),
 * x_aug [n_aug*B,N,2] (n_aug 1 or 8; may be NULL).                                       */
int htsp_extract_normalize(const float* data, const int64_t* fragment, int B, int Ntot, int N,
                           int n_aug, float* x_new, float* scale, float* x_aug, void* stream);
int htsp_augment8(const float* x, int B, int N, float* x_aug, void* stream);

/* ---- K2: encoder --------------------------------------------------------------------
 * MHAEncoder.forward, eval mode (models.py:30-60; utils.py:12-34): x [Bp,N,2] -> emb [Bp,N,128]. */
size_t htsp_encode_workspace_bytes(int Bp, int N, int mode);
int htsp_encode(const void* blob, int n_layers, const float* x, int Bp, int N, int mode,
                float* emb, void* workspace, size_t workspace_bytes, void* stream);

/* MHAEncoder.forward with the BatchNorm1d layers in TRAIN mode (models.py:35,37 as PathSolver.training_step runs the
 * encoder, rl4cop/train_path_solver.py:352-375): every norm layer normalises with the mean and biased variance of the
 * batch (all Bp*N rows) and its own gamma / beta.  bn_gamma_beta: device [n_layers][4][128] f32 = norm1.weight |
 * norm1.bias | norm2.weight | norm2.bias per layer; bn_batch_stats: device [n_layers][4][128] f32 out = batch mean |
 * biased variance of norm1, then of norm2 (the caller updates running_mean / running_var with its momentum; the
 * unbiased variance is var * M / (M - 1), M = Bp*N).  Tensor-core modes only (HTSP_MODE_TC_X3 / _FAST), else
 * HTSP_ERR_UNSUPPORTED; deterministic (fixed-order fp64 reductions).  The running statistics inside `blob` are not used. */
size_t htsp_encode_train_workspace_bytes(int Bp, int N, int mode);
int htsp_encode_train(const void* blob, int n_layers, const float* bn_gamma_beta, const float* x, int Bp, int N,
                      int mode, float* emb, float* bn_batch_stats, void* workspace, size_t workspace_bytes,
                      void* stream);

/* ---- K3: POMO rollout ---------------------------------------------------------------
 * PathDecoder.reset/forward (models.py:360-446), GroupState/Env (train_path_solver.py:23-149)
 * and the decode loop + pi extraction (train_path_solver.py:252-288), greedy.
 * src,tgt [Bp] i32; first [G] i32 (the reference's randperm(N)[:G], :256, drawn by the host);
 * forced_pi [Bp,G,N] i32 or NULL (teacher forcing); outputs pi [Bp,G,N] i32 (cycle order
 * from first[g]) and reward [Bp,G] f32 (= -(cycle length - |x[src]-x[tgt]|));
 * logits_out [Bp,G,N-2,N] f32 or NULL receives the masked 10*tanh logits of every step.  */
size_t htsp_decode_workspace_bytes(int Bp, int N, int G, int mode);
int htsp_decode(const void* blob, int n_layers, const float* x, const float* emb, const int32_t* src,
                const int32_t* tgt, const int32_t* first, int Bp, int N, int G, int mode,
                const int32_t* forced_pi, int32_t* pi, float* reward, float* logits_out,
                void* workspace, size_t workspace_bytes, void* stream);

/* Sampling rollout (forward only): PathSolver.forward(greedy=False),
 * train_path_solver.py:266-271 -- every row draws its next node from softmax(masked 10*tanh logits)
 * (`probs.multinomial(1)`) instead of the arg-max.  Drawn in the kernel by the Gumbel-max rule with
 * Philox4x32-10 noise keyed by (seed; instance, row, step, key): the tours depend on `seed` only, not on the
 * launch geometry; they match the reference in distribution, never bit-wise (different generator).
 * Served by the tcgen05 kernel only (tensor-core mode, 17 <= N <= 208, G <= 256), else HTSP_ERR_UNSUPPORTED. */
int htsp_decode_sample(const void* blob, int n_layers, const float* x, const float* emb,
                       const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                       int G, int mode, uint64_t seed, int32_t* pi, float* reward, void* workspace,
                       size_t workspace_bytes, void* stream);

/* As htsp_decode_sample, additionally returning logp [Bp,G] f32 = sum over the N-2 decode steps of
 * log p(sampled node), p = softmax(masked 10*tanh logits): the `group_prob_list.log().sum(dim=2)` of the
 * reference's training step (train_path_solver.py:394-414), forward only.
 * Served by the second tcgen05 kernel (HTSP_MODE_TC_FAST; 33 <= N <= 208 with G <= 256, or 208 < N <= 576 in key
 * blocks), else HTSP_ERR_UNSUPPORTED. */
int htsp_decode_sample_logp(const void* blob, int n_layers, const float* x, const float* emb,
                            const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                            int G, int mode, uint64_t seed, int32_t* pi, float* reward, float* logp,
                            void* workspace, size_t workspace_bytes, void* stream);

/* How many CTAs of the rollout kernel share an SM for an htsp_decode call of this shape on the current device: 2 when
 * mode HTSP_MODE_TC_FAST runs its single-tile CTAs (more CTAs than SMs; HTSP_T2_ONE = 0 | 1 overrides), else 1; 0 for
 * an invalid shape.  The result does not depend on it (bit-identical); bench.py and the tests report it.          */
int htsp_decode_ctas_per_sm(int Bp, int N, int G, int mode);

/* Operand-stream geometry of the HTSP_MODE_TC_FAST rollout kernel for subproblems of N nodes (host logic, no CUDA call;
 * tests/test_abi.py checks it on the CPU): out5 = { keys per key block (a multiple of 16; <= 208 with one block, <= 192
 * with several), key blocks (1 up to N = 208, 2 or 3 up to N = 576), ring stages per decode step, bytes of the operand
 * image of one instance, bytes of a ring slot }.  HTSP_ERR_UNSUPPORTED outside 33 <= N <= 576.                      */
int htsp_decode_key_blocks(int N, int32_t* out5);
/* Stage i (0 <= i < stages per decode step) of that stream: out2 = { byte offset into the instance's operand image,
 * bytes }.  The stages of a step tile the image exactly once and each fits a ring slot.                            */
int htsp_decode_key_block_stage(int N, int i, uint32_t* out2);

/* Test hook (tests/test_parity_gpu.py): as htsp_decode in a tensor-core mode on the tcgen05 kernel
 * (N <= 224, G <= 256), additionally dumping the intermediates of instance 0 at decode step
 * `dbg_step` into dbg [htsp_decode_debug_floats(N)] f32: raw glimpse scores S [2,8,128,NP]
 * (row tile, head, row, key; base-2 exponent units), normalised glimpse outputs O [256,128] and raw
 * pointer scores U [256,NP] (before c0 and 1/sqrt(128)); NP = N rounded up to 16.              */
size_t htsp_decode_debug_floats(int N);
int htsp_decode_debug(const void* blob, int n_layers, const float* x, const float* emb,
                      const int32_t* src, const int32_t* tgt, const int32_t* first, int Bp, int N,
                      int G, int mode, const int32_t* forced_pi, int32_t* pi, float* reward,
                      void* workspace, size_t workspace_bytes, float* dbg, int dbg_step, void* stream);

/* ---- K4: best-of selection and path orientation ---------------------------------------
 * train_path_solver.py:290-306 (max over G, then over the 8 augmentation blocks; first max
 * wins) -> best_len [B] (= -max reward), best_pi [B,N] i64, best_row [B] i32 (a*G+g).
 * htsp_finalize_paths: h_tsp.py:206-226 -- lengths / s, rotate to start at local 0, flip if
 * it does not end at N-1, map through fragment to global ids; status[b] != 0 flags a row
 * that violates the reference's asserts (h_tsp.py:218-224).                              */
int htsp_select_best(const float* reward, const int32_t* pi, int n_aug, int B, int G, int N,
                     float* best_len, int32_t* best_row, int64_t* best_pi, void* stream);
int htsp_finalize_paths(const int64_t* best_pi, const int64_t* fragment, const float* best_len,
                        const float* scale, int B, int N, int64_t* out_paths, float* out_len,
                        int32_t* status, void* stream);

/* ---- K5: tour length ----------------------------------------------------------------
 * rl_utils.get_tour_distance over dist_matrix = fp32(cdist_fp64(1000 x)) (rl_utils.py:87-102,
 * h_tsp.py:449-453,586-589) without the N x N matrix: closed-tour length of each of E tours.
 * coords [E,Ntot,2] f32; tours [E,Tmax] i64; tour_len [E] i32 -> out [E] f32 (already / 1000). */
int htsp_tour_length(const float* coords, const int64_t* tours, const int32_t* tour_len, int E,
                     int Ntot, int Tmax, float* out, void* stream);

/* ---- K6: large-TSP environment state ------------------------
 * Device-resident replacement of the per-step bookkeeping of Env._step (h_tsp.py:569-603), batched
 * over E environments (VecEnv.step runs them in a Python loop, h_tsp.py:779-782).
 * htsp_env_splice = LargeState.move_to's list splice (h_tsp.py:288-320): new_paths [E,P] i64 with
 * path_len [E] i32 valid entries each; tour_in/tour_out [E,Ntot] i64 (distinct buffers), tour_len [E]
 * i32 in/out, pos [E,Ntot] i32 in/out = position of every city in its tour or -1; status [E] i32:
 * 0 ok, 1 an endpoint of the path is not on the tour, 2 both endpoints coincide, 3 duplicate city,
 * 4 the tour did not grow, 5 index out of range (the reference's asserts, :293-320).
 * htsp_env_update (after the splice, on the new tour): neighbour coordinates
 * (rl_utils.update_neighbor_coord_, rl_utils.py:135-159), selected / available masks with the k-NN
 * lists knn [E,Ntot,k] i64 (LargeState.mask, h_tsp.py:330-340), the closed-tour length
 * (get_tour_distance over fp32(fp64 cdist x 1000), rl_utils.py:87-102, without the Ntot x Ntot matrix;
 * cur_len [E] in: old length, out: new length / 1000), reward [E] = old - new (h_tsp.py:599), and the
 * [E,Ntot,8] state tensor of LargeState.to_tensor (h_tsp.py:396-408; may be NULL).             */
int htsp_env_splice(const int64_t* new_paths, const int32_t* path_len, int E, int P, int Ntot,
                    const int64_t* tour_in, int64_t* tour_out, int32_t* tour_len, int32_t* pos,
                    int32_t* status, void* stream);
int htsp_env_update(const float* x, const int64_t* tour, const int32_t* tour_len, const int64_t* new_paths,
                    const int32_t* path_len, const int64_t* knn, int k, int E, int P, int Ntot,
                    uint8_t* selected, uint8_t* available, float* neighbor_coord, float* state,
                    float* cur_len, float* reward, void* stream);

/* ---- K7: fragment selection ----------------------------------
 * htsp_env_knn = the k-NN lists of Env.reset (h_tsp.py:443-455): for every city the k nearest
 * other cities under float32(float64 distance x 1000), ascending, equal distances by city id;
 * x [E,Ntot,2] f32 -> knn [E,Ntot,k] i64.  No Ntot x Ntot matrix is materialised.
 * htsp_env_fragment = Env.get_fragment_knn (h_tsp.py:482-542) for E environments: action [E,2] f32
 * is the policy output after Env.scale_action ([0,1]^2); selected [E,Ntot] u8, tour / tour_len /
 * pos as in K6; active [E] u8 or NULL.  Outputs: fragment [E,frag_len] i64 (what RLSolver.solve
 * receives), new_cities [E,frag_len] i64 (-1 padded) with n_new [E], start_city [E] i32 (-1 for an
 * empty tour), status [E] (1 = one of the reference's fragment-length asserts, h_tsp.py:641,656-658). */
int htsp_env_knn(const float* x, int E, int Ntot, int k, int64_t* knn, void* stream);
int htsp_env_fragment(const float* x, const uint8_t* selected, const int64_t* knn, int k,
                      const int64_t* tour, const int32_t* tour_len, const int32_t* pos,
                      const float* action, const uint8_t* active, int E, int Ntot, int frag_len,
                      int max_new_nodes, int64_t* fragment, int64_t* new_cities, int32_t* n_new,
                      int32_t* start_city, int32_t* status, void* stream);

/* ---- K8: upper-level policy forward ---------------------------
 * HTSP_PPO.forward (h_tsp.py:1272-1276): ActorPPO.forward (rl_models.py:329-336) o IMPALAEncoder.forward
 * (rl_models.py:187-273) o IMPALACNN.forward (rl_models.py:117-133) with the reference's default geometry
 * (input_dim 8, bev grid 100 x 100 over [0,1]^2, depths [16,32,32]); embed = cfg.embedding_dim (128),
 * mid = cfg.acotr_mid_dim (128), adim = cfg.nb_actions (2).
 * htsp_policy_pack_weights: host fp32 tensors in the order of HTSP_PPO.state_dict() restricted to
 *   encoder.input_transform.{0.weight,1.weight,1.bias}; per block i = 0..2: encoder.cnn.feat_convs.i.0,
 *   resnet1.i.1, resnet1.i.3, resnet2.i.1, resnet2.i.3 (weight, bias each); encoder.cnn.fc;
 *   actor.net.0.0, actor.net.0.2, actor.net.1, actor.net.3 (weight, bias each)  -> device blob.
 * htsp_policy_forward: state [E,Ntot,8] f32 (the tensor htsp_env_update writes) -> feat [E,embed]
 *   (graph_feat) and action [E,adim] = tanh(actor.net(feat)); image_out (may be NULL) receives the pseudo
 *   image [E,embed/2,100,100] the reference hands to its CNN.  Pillar ids must stay exact in fp32
 *   (E * 10^4 < 2^24, rl_models.py:207-211).                                                     */
int htsp_policy_n_tensors(void);
size_t htsp_policy_blob_bytes(int embed, int mid, int adim);
int htsp_policy_pack_weights(const float* const* host_tensors, int n_tensors, int embed, int mid,
                             int adim, void* dev_blob, void* stream);
size_t htsp_policy_workspace_bytes(int E, int Ntot, int embed);
int htsp_policy_forward(const void* blob, int embed, int mid, int adim, const float* state, int E,
                        int Ntot, float* feat, float* action, float* image_out, void* workspace,
                        size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* HTSP_B200_H_ */
