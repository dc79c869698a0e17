/*
 * mmae_b200.h -- C ABI of libmmae_b200.so, the B200 (sm_100a) engine for the MMAE hot path.
 *
 * The reference (natashamjaques/MultimodalAutoencoder) has no FFI: its only runtime
 * boundary is tf.Session.run(fetches, feed_dict) on the graph built in
 * multimodal_autoencoder.py:344-452.  Each entry point below replaces one family of
 * session.run call shapes (cited per function); the Python class
 * multimodalautoencoder_b200.MultimodalAutoencoder binds them through ctypes and keeps
 * the reference's own method names on top (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - every function returns 0 on success, a negative mmae_status otherwise; the text is
 *    available from mmae_last_error(engine) (or mmae_last_error(NULL) for create failures);
 *  - all matrices are fp32, row-major [batch, features]; weights are [in, out] (y = x.W + b,
 *    multimodal_autoencoder.py:467);
 *  - pointers named *_dev are device pointers on the engine's device, *_host are host pointers;
 *  - calls are asynchronous on the engine's stream unless stated; an engine is not thread safe.
 */
#ifndef MMAE_B200_H
#define MMAE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmae_engine mmae_engine;

typedef enum {
  MMAE_OK = 0,
  MMAE_ERR_INVALID = -1,   /* bad argument / unsupported configuration   */
  MMAE_ERR_CUDA = -2,      /* CUDA runtime or driver error (sticky)       */
  MMAE_ERR_NOTFOUND = -3,  /* unknown variable name                       */
  MMAE_ERR_STATE = -4,     /* call order (e.g. head step without a head)  */
  MMAE_ERR_COMM = -5       /* NCCL unavailable or failed                  */
} mmae_status;

/* multimodal_autoencoder.py:477-497 */
typedef enum { MMAE_ACT_LINEAR = 0, MMAE_ACT_RELU = 1, MMAE_ACT_TANH = 2,
               MMAE_ACT_SOFTSIGN = 3, MMAE_ACT_SOFTPLUS = 4 } mmae_activation;
/* multimodal_autoencoder.py:381-390 */
typedef enum { MMAE_LOSS_RMSE = 0,            /* 'mean_squared' (really an RMSE, :383-384) */
               MMAE_LOSS_SIGMOID_CE = 1,      /* batch SUM, :388                           */
               MMAE_LOSS_CE = 2 } mmae_loss;  /* -sum(x log xhat), :386                    */
/* multimodal_autoencoder.py:431-438 */
typedef enum { MMAE_HEAD_SIGMOID_CE = 0, MMAE_HEAD_SOFTMAX_CE = 1 } mmae_head_loss;
typedef enum { MMAE_PREC_FP32 = 0,            /* CUDA-core fp32 FMA everywhere             */
               MMAE_PREC_TF32 = 1 } mmae_precision; /* tcgen05 kind::tf32 where shapes allow */
typedef enum { MMAE_NOISE_INTELLIGENT = 0,    /* categorical over noise types, :686-695    */
               MMAE_NOISE_UNIFORM = 1 } mmae_noise_mode; /* randint modality, :698-700     */

/* Constructor arguments of MultimodalAutoencoder (multimodal_autoencoder.py:59-70) that reach
 * the graph, after the ctor's own overrides (:175-184). */
typedef struct {
  int32_t num_feats;                 /* F = data_loader.num_feats                           */
  int32_t num_modalities;            /* M <= 32                                             */
  const int32_t* modality_starts;    /* [M+1], last == F  (data_funcs.py:121-122)           */
  int32_t num_layers;                /* L = len(layer_sizes)                                */
  const int32_t* layer_sizes;        /* [L]                                                 */
  int32_t tie_weights;
  int32_t variational;
  int32_t activation;                /* mmae_activation                                     */
  int32_t loss_func;                 /* mmae_loss                                           */
  float weight_penalty;
  float learning_rate;               /* constant: the reference's decay is inert (:356-361 vs :411) */
  float beta1, beta2, adam_eps;      /* 0.9, 0.999, 1e-8 = tf.train.AdamOptimizer defaults  */
  int32_t num_head_layers;           /* len(classification_layer_sizes)+1, 0 = no head      */
  const int32_t* head_sizes;         /* [num_head_layers] output widths, last = #logits     */
  int32_t head_activation;
  int32_t head_loss;                 /* mmae_head_loss                                      */
  float head_weight_penalty;
  float head_learning_rate;
  float mask_with;                   /* -1.0                                                */
  int32_t n_zero;                    /* int(F * .05), :682                                  */
  int32_t noise_mode;                /* mmae_noise_mode                                     */
  int32_t num_noise_types;           /* K                                                   */
  const uint32_t* noise_type_masks;  /* [K] modality bitmask per noise type (:203-206)      */
  const uint32_t* noise_thresholds;  /* [K-1] cumulative P as floor(c * 2^32) (:202)        */
  int32_t num_modalities_to_drop;    /* <= 4                                                */
  uint64_t seed;                     /* Philox key                                          */
  int32_t precision;                 /* mmae_precision                                      */
  int64_t max_batch;                 /* workspace hint; grown on demand                     */
  /* comparison_algorithms/neural_net.py on the same kernels: a plain MLP classifier = the encoder stack
   * with EVERY layer activated (+ dropout) followed by one linear logits layer (the head, num_head_layers == 1), no
   * decoder in the loss; trained by the head optimizer with head_weight_penalty on every weight matrix (:181-182). */
  int32_t classifier_only;
  float clip_norm;                   /* > 0: tf.clip_by_global_norm(gradients, clip_norm) before Adam (:189-190)  */
} mmae_config;

/* Bits of `want` for mmae_forward */
#define MMAE_WANT_RECON     1u   /* decoded_X           (:378/:390)  -> out->recon      [B,F]  */
#define MMAE_WANT_EMBEDDING 2u   /* embedding           (:367/:375)  -> out->embedding  [B,E]  */
#define MMAE_WANT_HEAD      4u   /* logits/probs/preds  (:428-450)   -> out->logits/probs/preds */
#define MMAE_WANT_LOSS      8u   /* reconstruction_loss (:381-390), needs target              */
#define MMAE_WANT_FILLED    16u  /* recon on missing blocks only (data_funcs.py:310-381)      */
#define MMAE_WANT_HEAD_LOSS 32u  /* classification_loss + accuracy (:431-452), needs labels   */

typedef struct {
  float* recon;      /* [B,F] device, or NULL */
  float* embedding;  /* [B,E] */
  float* logits;     /* [B,C] */
  float* probs;      /* [B,C] sigmoid(logits), :446 */
  int32_t* preds;    /* [B,C] (sigmoid head) or [B] (softmax head) */
  float* filled;     /* [B,F] */
} mmae_outputs;

/* Scalars of the last call, read with mmae_read_scalars (synchronises the stream). */
typedef enum {
  MMAE_S_RECON_LOSS = 0,  /* reconstruction_loss as the graph defines it (sum for the CE losses) */
  MMAE_S_KL_MEAN = 1,     /* mean_b KL_b (:402-406), 0 when not variational                       */
  MMAE_S_SUMSQ = 2,       /* sum (xhat-x)^2 (RMSE loss only)                                      */
  MMAE_S_HEAD_LOSS = 3,   /* classification data loss (mean CE, without the L2 term)              */
  MMAE_S_HEAD_ACC = 4,    /* accuracy (:451-452)                                                  */
  MMAE_S_GRAD_SCALE = 5,  /* 1/(N*rmse) applied to the RMSE gradient inside Adam                  */
  MMAE_NUM_SCALARS = 8
} mmae_scalar;

/* ---- lifetime: tf.Graph + tf.Session + global_variables_initializer (:233-237, :542-547) ---- */
int mmae_create(const mmae_config* cfg, mmae_engine** out);
void mmae_destroy(mmae_engine* e);
const char* mmae_last_error(const mmae_engine* e);
int mmae_set_stream(mmae_engine* e, void* cuda_stream);
int mmae_synchronize(mmae_engine* e);

/* ---- variables: names are the reference's tf.Variable names (:279-334); used by
 *      initialisation and by save_model/load_saved_model (:766-859) ---- */
int mmae_num_variables(const mmae_engine* e);
int mmae_variable_info(const mmae_engine* e, int index, char* name_out, int name_cap,
                       int64_t* rows, int64_t* cols);
int mmae_set_variable(mmae_engine* e, const char* name, const float* host, int64_t count);
int mmae_get_variable(mmae_engine* e, const char* name, float* host, int64_t count);
int mmae_get_gradient(mmae_engine* e, const char* name, float* host, int64_t count);
/* optimizer 0 = opt_step (:411), 1 = classification_opt_step (:443); slots m, v and step t */
int mmae_get_opt_state(mmae_engine* e, int optimizer, const char* name, float* m_host, float* v_host,
                       int64_t count, int64_t* t);
int mmae_set_opt_state(mmae_engine* e, int optimizer, const char* name, const float* m_host,
                       const float* v_host, int64_t count, int64_t t);

/* ---- block-mask noise: add_noise_to_batch + mask_modality (:649-702) ----
 * The descriptor is (zero bitmap [B, ceil(F/32)] uint32, modality bitmask [B] uint32).
 * mmae_set_noise uploads one built on the host in the reference's NumPy call order
 * (rng_mode='numpy'); mmae_gen_noise draws it on the device from Philox (rng_mode='philox').
 * Kernels apply it while loading X; mmae_apply_noise materialises noisy_X for callers
 * that want the array itself. */
int mmae_set_rng_step(mmae_engine* e, uint64_t step);
int mmae_set_noise(mmae_engine* e, const uint32_t* zero_bits_host, const uint32_t* mod_bits_host,
                   int64_t batch);
int mmae_gen_noise(mmae_engine* e, int64_t batch, int64_t first_row);
int mmae_get_noise(mmae_engine* e, uint32_t* zero_bits_host, uint32_t* mod_bits_host, int64_t batch);
int mmae_apply_noise(mmae_engine* e, const float* X_dev, int64_t batch, float* out_dev);

/* ---- forward-only fetches: session.run of decoded_X / reconstruction_loss / embedding /
 *      predictions / [classification_loss, accuracy] (:726-730, :759-762, :945, :1013, :1029,
 *      :1044, :1078, :1123, :1158) ----
 * use_noise != 0 applies the current descriptor to X_dev on load (noisy_X feed); target_dev
 * is the true_X feed (NULL = X_dev itself, as predict() does at :941-942); labels_dev is the
 * true_Y feed.  keep is the tf_dropout_prob feed. */
int mmae_forward(mmae_engine* e, const float* X_dev, const float* target_dev, const float* labels_dev,
                 int64_t batch, int use_noise, float keep, uint32_t want, const mmae_outputs* out);

/* ---- session.run([opt_step]) (:590): forward + backward + Adam on the autoencoder ----
 * use_noise: 0 = X_dev is fed as it is; 1 = apply the descriptor loaded by mmae_set_noise / mmae_gen_noise;
 * MMAE_NOISE_DRAW (3) = draw this step's Philox descriptor inside the call (add_noise_to_batch, :668-702, fused with
 * the production of the noisy batch: one kernel instead of a draw pass and an apply pass). */
#define MMAE_NOISE_DRAW 3
int mmae_train_step(mmae_engine* e, const float* X_dev, int64_t batch, int use_noise, float keep);
/* Same step with the two feeds given separately, as feed_dict {noisy_X: ..., true_X: ...} does (:570-571):
 * X_in_dev is what the encoder reads (an already-noised matrix when use_noise == 0), target_dev what the
 * loss compares against. */
int mmae_train_step_pair(mmae_engine* e, const float* X_in_dev, const float* target_dev, int64_t batch,
                         int use_noise, float keep);
/* ---- session.run([classification_opt_step]) (:647) ---- */
int mmae_cls_train_step(mmae_engine* e, const float* X_dev, const float* labels_dev, int64_t batch,
                        int use_noise, float keep);
/* Same two steps fed from HOST buffers (what feed_dict does): pinned-or-pageable fp32 in,
 * H2D inside the call.  gen_noise: 0 = clean input, 1 = Philox noise drawn on the device for this step,
 * 2 = apply the descriptor last given to mmae_set_noise (host RNG, reference order). */
int mmae_train_step_host(mmae_engine* e, const float* X_host, int64_t batch, int gen_noise, float keep);
int mmae_cls_train_step_host(mmae_engine* e, const float* X_host, const float* labels_host,
                             int64_t batch, int gen_noise, float keep);
/* Forward fed from / returning to HOST buffers (predict(), :932-950). `out` holds host pointers. */
int mmae_forward_host(mmae_engine* e, const float* X_host, const float* target_host,
                      const float* labels_host, int64_t batch, int use_noise, float keep,
                      uint32_t want, const mmae_outputs* out_host);

/* Split step for data-parallel callers that own the collective: backward leaves the UNSCALED
 * flat gradient (+ loss partial sums in the tail) in the buffer returned by mmae_grad_buffer;
 * after an external sum-allreduce of that buffer, mmae_apply_update runs the fused
 * scale + L2 + Adam.  mmae_train_step == backward + (engine NCCL allreduce if a communicator
 * is attached) + apply_update. */
int mmae_backward(mmae_engine* e, const float* X_dev, int64_t batch, int64_t global_batch,
                  int use_noise, float keep);
int mmae_grad_buffer(mmae_engine* e, float** dev_ptr, int64_t* count);
int mmae_apply_update(mmae_engine* e, int optimizer);

/* ---- device-resident dataset + on-device batch sampling (data_funcs.py:161-195) ---- */
int mmae_set_dataset(mmae_engine* e, int slot, const float* X_host, const float* Y_host,
                     int64_t rows, int32_t label_cols);
/* mmae_set_dataset with the rows already in device memory (device-to-device copy into engine-owned buffers). */
int mmae_set_dataset_device(mmae_engine* e, int slot, const float* X_dev, const float* Y_dev, int64_t rows, int32_t label_cols);

/* A training VIEW of a resident dataset: the rows a cross-validation fold trains on (set_to_cross_validation_fold,
 * data_funcs.py:278-308) as a list of dataset rows.  The dataset is uploaded once; switching folds uploads count indices
 * instead of the matrix.  Sampled / given indices then address the view (index j = dataset row rows_host[j]).
 * rows_host == NULL clears the view. */
int mmae_set_dataset_view(mmae_engine* e, int slot, const int64_t* rows_host, int64_t count);

/* idx_host == NULL draws rows from Philox; otherwise the caller's indices (np.random.choice). */
int mmae_train_step_resident(mmae_engine* e, int slot, const int64_t* idx_host, int64_t batch,
                             int gen_noise, float keep, int classification);

/* reconstruction_loss (:726) of a batch sampled on the device from a resident dataset (rows, noise and dropout as in
 * mmae_train_step_resident, no update): the train-loss fetch of a record step without a host round trip. */
int mmae_eval_resident(mmae_engine* e, int slot, int64_t batch, int gen_noise, float keep);

/* ---- get_reconstruction_loss_per_modality (:1189-1216) as one batched pass: for every modality m the rows are
 *      reconstructed with block m set to the literal -1.0 (:1203) and rmse_host[m] receives sqrt(mean((X - X_hat)^2)) over
 *      that block's columns.  The M masked copies are stacked into one batch per chunk of rows (one forward for all
 *      modalities), the squared errors are reduced on the device; X_host is [rows, num_feats] fp32. ---- */
int mmae_modality_rmse(mmae_engine* e, const float* X_host, int64_t rows, double* rmse_host);

/* ---- scalars of the last call ---- */
int mmae_read_scalars(mmae_engine* e, double* out, int count);

/* Same, without synchronising: enqueues the device->host copy on the engine's stream into caller-owned
 * PINNED memory (the per-step loss read of a training loop that must not stall the pipeline). */
int mmae_read_scalars_async(mmae_engine* e, double* pinned_host, int count);

/* ---- data parallel: engine-owned NCCL communicator (libnccl is dlopen'ed on first use) ---- */
int mmae_comm_unique_id(void* id_out_128);
int mmae_comm_init(mmae_engine* e, const void* id_128, int rank, int world_size);
/* Shard description for data-parallel runs: loss normalisers (RMSE N, KL 1/B, head mean) use
 * global_batch, and the Philox streams are indexed by first_row + local row so that G ranks
 * reproduce the 1-rank masks bit for bit.  global_batch = 0 means "the local batch". */
int mmae_set_shard(mmae_engine* e, int64_t global_batch, int64_t first_row);

/* ---- introspection used by tests and bench ---- */
int64_t mmae_kernel_launches(const mmae_engine* e);   /* kernels launched since create */
int64_t mmae_chain_launches(const mmae_engine* e);    /* of those, whole-network (encode+decode+loss) launches */
int64_t mmae_backward_chain_launches(const mmae_engine* e);   /* ... and whole-backward (every dgrad of the step) launches */
int64_t mmae_wgrad_group_launches(const mmae_engine* e);      /* ... and grouped weight-gradient (every dW of the step) launches */
int64_t mmae_graph_replays(const mmae_engine* e);     /* train steps replayed from a captured CUDA graph (MMAE_GRAPHS=0 disables) */
int64_t mmae_fused_noise_launches(const mmae_engine* e); /* GEMM launches that applied mask + noise in their A-operand load (:668-702) */
/* Device-side timing of the tcgen05 GEMM launches (CUDA events on the engine's stream around each
 * launch); read returns the accumulated milliseconds, algorithmic FLOPs (2*M*N*K) and launch count
 * since profiling was switched on. */
int mmae_set_profiling(mmae_engine* e, int on);
int mmae_read_profile(mmae_engine* e, double* gemm_ms, double* gemm_flops, int64_t* gemm_launches);
int mmae_get_buffer(mmae_engine* e, const char* name, float* host, int64_t count); /* "eps","mu","lv","emb","out","logits" */
/* Inject the VAE epsilon (tf.random_normal, :374) for parity runs; NULL returns to Philox draws. */
int mmae_set_eps(mmae_engine* e, const float* eps_host, int64_t count);
/* C = op(A) op(B) (+bias) through the engine's GEMM families; precision as mmae_precision.
 * transA: A stored [K,M]; transB: B stored [N,K].  All device pointers. */
int mmae_debug_gemm(int precision, int transA, int transB, int64_t M, int64_t N, int64_t K,
                    const float* A_dev, int64_t lda, const float* B_dev, int64_t ldb,
                    float* C_dev, int64_t ldc, const float* bias_dev, int activation, float beta,
                    void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* MMAE_B200_H */
