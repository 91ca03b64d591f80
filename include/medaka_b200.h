/*
 * medaka_b200.h - C ABI of libmedaka_b200.so, the sm_100a engine behind medaka's
 * inference hot path (medaka/prediction.py:44-52).
 *
 * Plain C: opaque handles, pointers + sizes, int return codes (0 = MDK_OK).  No torch
 * types, no exit(): every failure returns a negative code and leaves a message in
 * mdk_last_error() (the reference's C layer calls exit(1) on error,
 * src/medaka_common.c:19-26; a drop-in library must not).
 *
 * Each entry point cites the reference interface it replaces.  The cffi binding a
 * medaka maintainer would add is in INTEGRATION.md; medaka_b200/libmedaka.py is
 * that binding (it cdef()s this file minus the '#' lines, like the reference's
 * build.py:71-82 does for libmedaka).
 *
 * Conventions
 *   - "host" pointers are ordinary host memory (pinned memory from mdk_host_alloc makes
 *     the copies asynchronous); "dev" pointers are CUDA device pointers on the engine's
 *     device (e.g. torch tensor .data_ptr()).
 *   - feats   float32 [B][T][F] row-major (torch_ext.Batch.counts_matrix, torch_ext.py:155)
 *   - probs   float32 [B][T][5]  (GRUModel.forward output, gru.py:58-72)
 *   - logits  float32 [B][T][5]  (pre-softmax, gru.py:67)            - optional (NULL)
 *   - labels  uint8   [B][T]     (argmax, first max wins, labels.py:1063) - optional
 *   - weights: torch state-dict layout (gate order r,z,n): w_ih [3H][in], w_hh [3H][H],
 *     b_ih [3H], b_hh [3H]; linear w [5][2H], b [5]   
 */
#ifndef MEDAKA_B200_H
#define MEDAKA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDK_OK 0
#define MDK_ERR_ARG -1
#define MDK_ERR_CUDA -2
#define MDK_ERR_STATE -3
#define MDK_ERR_UNSUPPORTED -4
#define MDK_ERR_NOMEM -5

/* precision modes of the GRU gate matmuls */
#define MDK_PREC_TC 0    /* tcgen05 tensor cores, fp16 hi/lo split operands (3 MMAs), fp32 accumulate */
#define MDK_PREC_FP32 1  /* CUDA-core fp32 FFMA path: validation / --full_precision */

/* which recurrent kernel the tensor-core path runs (mdk_engine_set_rec_mode) */
#define MDK_REC_AUTO 0      /* ping-pong from half a wave of 16-window tiles up, else one tile per CTA */
#define MDK_REC_ONE_TILE 1  /* rec_tc_kernel: one (two beyond a wave) 16-window tile per CTA */
#define MDK_REC_PINGPONG 2  /* rec_pp_kernel: two tiles per CTA, their MMA and gate phases interleaved */

/* count normalisation modes: CountsFeatureEncoder._norm_modes_ (medaka/features.py:816) */
#define MDK_NORM_TOTAL 0
#define MDK_NORM_FWD_REV 1
#define MDK_NORM_NONE 2

typedef struct mdk_engine mdk_engine;
typedef struct mdk_bam mdk_bam;              /* an open BAM file (+ its .bai index when present) */
typedef struct mdk_bam_batch mdk_bam_batch;  /* the records of one region, in BAM's packed encodings */

/* GRUModel constructor arguments (medaka/architectures/gru.py:13-21). */
typedef struct mdk_model_desc {
    int32_t num_features;   /* F = 10 * len(dtypes) */
    int32_t gru_size;       /* H; this build supports 128 (every shipped counts-matrix model) */
    int32_t n_layers;       /* 2 */
    int32_t bidirectional;  /* 1 */
    int32_t num_classes;    /* linear head is hard-coded to 5 outputs (gru.py:53-55) */
} mdk_model_desc;

/* per-stage device timings of the last mdk_engine_forward* call, milliseconds (CUDA events) */
typedef struct mdk_timings {
    float h2d_ms;
    float inproj0_ms;   /* layer-0 input projection */
    float rec0_ms;      /* layer-0 recurrence (both directions) */
    float inproj1_ms;   /* layer-1 input projection GEMM */
    float rec1_ms;      /* layer-1 recurrence */
    float head_ms;      /* linear + softmax + argmax */
    float d2h_ms;
    float total_ms;
    int32_t launches;   /* kernels launched by the call */
} mdk_timings;

/* constants the Python layer reads from the library, like libmedaka.lib.plp_bases etc.
 * (src/medaka_counts.h:19-22; read at medaka/common.py:29-35, medaka/features.py:860,902-904) */
const char *mdk_plp_bases(void);   /* "acgtACGTdD" */
size_t mdk_featlen(void);          /* 10 */
size_t mdk_fwd_del(void);          /* 9 */
size_t mdk_rev_del(void);          /* 8 */

const char *mdk_last_error(void);  /* thread-local message of the last failing call */
const char *mdk_version(void);
int mdk_device_count(int *count);
/* sm major*10+minor of `device`, SM count, bytes of global memory */
int mdk_device_info(int device, int *sm_arch, int *sm_count, size_t *total_mem);

/* pinned host memory for batch staging (replaces the pageable .to(device) / .cpu() copies of
 * TorchModel.predict_on_batch, medaka/models.py:309,312) */
int mdk_host_alloc(size_t bytes, void **out);
int mdk_host_free(void *p);
/* raw device memory + copies, for callers that keep inputs resident in HBM (bench, tests) */
int mdk_dev_alloc(int device, size_t bytes, void **out);
int mdk_dev_free(int device, void *p);
int mdk_memcpy_h2d(int device, void *dst_dev, const void *src_host, size_t bytes);
int mdk_memcpy_d2h(int device, void *dst_host, const void *src_dev, size_t bytes);
int mdk_dev_memset(int device, void *dst_dev, int value, size_t bytes);
int mdk_device_synchronize(int device);

/* ---- model seam: replaces ModelStoreTGZ.load_model + TorchModel.predict_on_batch --------------
 * (medaka/datastore.py:135-157, medaka/models.py:303-313) */
int mdk_engine_create(int device, const mdk_model_desc *desc, mdk_engine **out);
int mdk_engine_destroy(mdk_engine *e);
/* load_state_dict for one (layer, direction) of gru.* ; direction 1 = "_reverse" */
int mdk_engine_load_gru(mdk_engine *e, int layer, int direction, const float *w_ih,
                        const float *w_hh, const float *b_ih, const float *b_hh);
int mdk_engine_load_linear(mdk_engine *e, const float *w, const float *b);
/* TorchModel.half() / --full_precision (medaka/prediction.py:164-168): MDK_PREC_* */
int mdk_engine_set_precision(mdk_engine *e, int mode);
int mdk_engine_get_precision(mdk_engine *e, int *mode);
/* fp16 products per tensor-core contraction, a bit set: 1 = W_hi.x_hi (required), 2 = W_hi.x_lo, 4 = W_lo.x_hi.
 * 7 (default) reproduces fp32 to ~2e-6; the 2- and 1-product sets trade parity for tensor time - measured in
 * profiles/precision_r02.md; they do NOT meet the labels-bit-exact bar and are never selected automatically. */
int mdk_engine_set_products(mdk_engine *e, int mask);
/* MDK_REC_*: recurrent-kernel selection (A/B measurements; AUTO is the default) */
int mdk_engine_set_rec_mode(mdk_engine *e, int mode);
/* pre-size the compute lanes (workspace + staging) for groups of up to B windows of T columns (otherwise grown on
 * demand, to the size of the batch that opens a group - i.e. without a reserve call nothing is coalesced) */
int mdk_engine_reserve(mdk_engine *e, int64_t B, int64_t T);
/* predict_on_batch with HOST buffers: H2D feats, forward, D2H probs (+logits, +labels when
 * non-NULL); returns when outputs are in host memory. */
int mdk_engine_forward(mdk_engine *e, const float *feats_host, int64_t B, int64_t T,
                       float *probs_host, float *logits_host, uint8_t *labels_host);
/* asynchronous form of the same call: returns once the batch is queued.  Host buffers must stay valid (and should be
 * page-locked) until mdk_engine_wait(ticket) returns.
 * Batches are COALESCED: consecutive submits with the same T collect in a group (their features are copied to the
 * device as they arrive, behind the previous group's compute) and the group runs as ONE forward over all of its
 * windows - the reference's default batches (medaka/prediction.py:14, 100-200 windows) are a fraction of what fills a
 * B200.  A group is launched when it is full (one wave of windows, mdk_engine_preferred_windows, or as far as the
 * buffers sized by mdk_engine_reserve reach), when a batch with another T arrives, or when somebody waits for one of
 * its tickets.  Groups alternate between two compute lanes (own stream, own workspace), so the layer-1 pass of one
 * group shares the GPU with the layer-0 pass of the next; small forwards (<= 2^18 positions: the B = 1 remainder
 * regions of prediction.py:196-209) rotate over 14 more lanes and run concurrently.  Results are identical to
 * uncoalesced forwards: windows never interact. */
int mdk_engine_submit(mdk_engine *e, const float *feats_host, int64_t B, int64_t T,
                      float *probs_host, float *logits_host, uint8_t *labels_host, int64_t *ticket);
int mdk_engine_wait(mdk_engine *e, int64_t ticket);
/* launch the group that is still collecting batches (if any) without waiting for it */
int mdk_engine_flush(mdk_engine *e);
/* most windows coalesced into one group; 0 (default) = one wave, 1 = never coalesce */
int mdk_engine_set_group_windows(mdk_engine *e, int64_t windows);
/* same forward with DEVICE buffers (inputs resident in HBM); asynchronous on the next compute lane (consecutive calls
 * alternate lanes and may overlap: give them distinct output buffers), complete after mdk_engine_sync(). */
int mdk_engine_forward_dev(mdk_engine *e, const float *feats_dev, int64_t B, int64_t T,
                           float *probs_dev, float *logits_dev, uint8_t *labels_dev);
int mdk_engine_sync(mdk_engine *e);
int mdk_engine_last_timings(mdk_engine *e, mdk_timings *out);
/* mean per-stage device times over the last n_last (<= 32) forward calls */
int mdk_engine_mean_timings(mdk_engine *e, int n_last, mdk_timings *out);
/* bracket a timed region on the engine stream with CUDA events (bench.py) */
int mdk_engine_timer_start(mdk_engine *e);
int mdk_engine_timer_stop(mdk_engine *e, float *elapsed_ms);
/* debugging / layer-wise parity: copy an internal activation of the last forward to host.
 * which: 0 = layer-0 output [B][T][2H] fp32, 1 = layer-1 output [B][T][2H] fp32 */
int mdk_engine_read_activation(mdk_engine *e, int which, float *out_host, int64_t n_floats);
/* keep != 0: leave the layer-1 output in HBM (for mdk_engine_read_activation(e, 1, ...)) by running the 5-class head as
 * its own kernel; default 0: the head is fused into the layer-1 recurrence whenever the batch is one tile per CTA */
int mdk_engine_keep_activations(mdk_engine *e, int keep);
/* number of kernels launched by this engine since creation (bench.py "gpu_launches") */
int64_t mdk_engine_launch_count(mdk_engine *e);
/* Number of windows per predict_on_batch call that fills the device exactly once: the recurrent kernel runs one CTA per
 * (16-window tile, direction), so 16 * (SMs / 2) windows = 1184 on a B200 is one full wave (the reference's --batch_size
 * default, medaka/prediction.py:14 / medaka.py, is sized for its own GPUs' memory; a 200-window batch uses 26 of 148 SMs).
 * Callers that own the batching (run_prediction) should coalesce to this size. */
int64_t mdk_engine_preferred_windows(mdk_engine *e);

/* ---- featuriser seam: replaces CountsFeatureEncoder._post_process_pileup --------------------
 * (medaka/features.py:871-935): depth = sum of counts, minor columns take the depth of their
 * major column (np.searchsorted side='left'), optional sym_indels fill, normalisation by
 * `mode`, float64 divide then cast to float32.  counts is what calculate_pileup returns
 * (size_t matrix, src/medaka_counts.h:5-14): uint64 [n][F], F = 10 * num_dtypes.
 * feats_out float32 [n][F]; depth_out int64 [n] (may be NULL).
 * Host-buffer version stages through the device; *_dev works on device pointers. */
int mdk_normalise_counts(int device, const uint64_t *counts, const int64_t *major,
                         const int64_t *minor, int64_t n, int32_t num_dtypes, int32_t mode,
                         int32_t sym_indels, float *feats_out, int64_t *depth_out);
int mdk_normalise_counts_dev(int device, const uint64_t *counts_dev, const int64_t *major_dev,
                             const int64_t *minor_dev, int64_t n, int32_t num_dtypes,
                             int32_t mode, int32_t sym_indels, float *feats_out_dev,
                             int64_t *depth_out_dev);

/* ---- featuriser seam, raw counts: replaces calculate_pileup (src/medaka_counts.c:199-372, declared
 * src/medaka_counts.h:105-108) for num_homop == 1.  htslib's BGZF/BAM decoding stays on the host
 * (medaka_b200/bam.py); the records arrive in BAM's own packed encodings:
 *   pos[n] 0-based reference start, flag[n], mapq[n], dtype[n] (datatype index, 0 when num_dtypes == 1),
 *   cigar[] uint32 (len << 4 | op) with cigar_off[n+1] (op index of each read's first op),
 *   seq[] 4-bit codes, two per byte, high nibble first, with seq_off[n+1] (byte offset of each read).
 * Region is [start, end) 0-based on one contig; reads failing the flag / min_mapQ filter of
 * src/medaka_bamiter.c:19-21 are skipped on the device (tag / RG filters are the reader's job).
 * Outputs like the plp_data struct (src/medaka_counts.h:5-14): counts uint64 [n_cols][10*num_dtypes] in
 * 'acgtACGTdD' order, major / minor [n_cols].  *n_cols_out is always set; if it exceeds max_cols the call
 * returns MDK_ERR_NOMEM without writing counts and the caller retries with larger buffers
 * (enlarge_plp_data, medaka_counts.c:266-271).  All pointers are HOST pointers. */
int mdk_pileup_counts(int device, int64_t n_rec, const int32_t *pos, const uint16_t *flag,
                      const uint8_t *mapq, const uint8_t *dtype, const uint32_t *cigar,
                      const int64_t *cigar_off, const uint8_t *seq, const int64_t *seq_off,
                      int32_t start, int32_t end, int32_t num_dtypes, int32_t min_mapq,
                      int64_t max_cols, uint64_t *counts_out, int64_t *major_out, int64_t *minor_out,
                      int64_t *n_cols_out);

/* calculate_pileup + _post_process_pileup in one device pass (src/medaka_counts.c:199-372 followed by
 * medaka/features.py:871-935): the records of mdk_pileup_counts in, the normalised float32 features [n_cols][10*num_dtypes],
 * depth [n_cols] (may be NULL) and positions out; the uint64 counts stay on the device.  mode / sym_indels as for
 * mdk_normalise_counts.  Normalising a region before it is split at coverage gaps (medaka/features.py:125-134) gives the
 * same numbers as normalising the pieces: an insertion column and its parent never sit on different sides of a gap.
 * *n_cols_out is always set; MDK_ERR_NOMEM (nothing written) when it exceeds max_cols. */
int mdk_pileup_features(int device, int64_t n_rec, const int32_t *pos, const uint16_t *flag,
                        const uint8_t *mapq, const uint8_t *dtype, const uint32_t *cigar,
                        const int64_t *cigar_off, const uint8_t *seq, const int64_t *seq_off,
                        int32_t start, int32_t end, int32_t num_dtypes, int32_t min_mapq, int32_t mode,
                        int32_t sym_indels, int64_t max_cols, float *feats_out, int64_t *depth_out,
                        int64_t *major_out, int64_t *minor_out, int64_t *n_cols_out);

/* ---- featuriser seam, read-level features: replaces calculate_read_alignment (src/medaka_read_matrix.c:277-615,
 * declared src/medaka_read_matrix.h:124-129) for one region of one contig.  Records as for mdk_pileup_counts plus
 *   qual[] base qualities (l_seq bytes per read, 0xff = absent) with qual_off[n+1],
 *   aux[] the raw optional fields with aux_off[n+1] (read for the `mv` move table -> dwell channel, calculate_dwells
 *   :154-213, and the `HP` haplotag, :405-414; may be NULL when neither channel is requested),
 *   names[] query names with name_off[n+1] (alignments of one name share a row, :386-388).
 * Output like read_aln_data (src/medaka_read_matrix.h:5-17) after the Python wrapper's `[:, :n_reads]` cut
 * (medaka/features.py:337-347): matrix int8 [n_cols][n_reads][featlen], featlen = 4 (+1 dwells) (+1 haplotype)
 * (+1 datatype when num_dtypes > 1): base 1..4 = ACGT / 5 = deletion / -1 other, base quality, strand +1 / -1, mapping
 * quality, ... ; cells no read touches are 0.  Values are NOT clipped (the wrapper's np.maximum(.., 0) is the caller's).
 * Rows follow the reference's bookkeeping (first row whose previous read ended >= 5 positions ago, or a new row;
 * row_per_read: always a new row; rows >= max_reads are dropped); n_reads = min(max_reads, deepest column) or the
 * number of rows used with row_per_read.  left_read_out / right_read_out [n_reads] (may both be NULL): index of the read
 * whose name the reference reports as read_ids_left / read_ids_right for that row, -1 = "__blank_<k>", -2 = NULL id.
 * *n_cols_out and *n_reads_out are always set; if n_cols > max_cols or n_cols * n_reads * featlen > max_cells the call
 * returns MDK_ERR_NOMEM without data and the caller retries with larger buffers (enlarge_read_aln_data_*, :85-138).
 * Known divergence: a read that finds no free row is dropped for good; the reference can let such a read alias a row
 * that is pushed after a later growth of its buffer (reads another read's struct - undefined behaviour, not reproduced).
 * All pointers are HOST pointers. */
int mdk_read_matrix(int device, int64_t n_rec, const int32_t *pos, const uint16_t *flag, const uint8_t *mapq,
                    const uint8_t *dtype, const uint32_t *cigar, const int64_t *cigar_off, const uint8_t *seq,
                    const int64_t *seq_off, const uint8_t *qual, const int64_t *qual_off, const uint8_t *aux,
                    const int64_t *aux_off, const char *names, const int64_t *name_off, int32_t start, int32_t end,
                    int32_t num_dtypes, int32_t min_mapq, int32_t row_per_read, int32_t include_dwells,
                    int32_t include_haplotype, int32_t max_reads, int64_t max_cols, int64_t max_cells,
                    int8_t *matrix_out, int64_t *major_out, int64_t *minor_out, int64_t *n_cols_out,
                    int32_t *n_reads_out, int32_t *left_read_out, int32_t *right_read_out);

/* ---- read-level model seam: LatentSpaceLSTM (medaka/architectures/latent_space_lstm.py:34-207, the model class of the
 * `rl_` consensus models) behind TorchModel.predict_on_batch (medaka/models.py:303-313) with
 * ReadLevelFeaturesModel.get_model_input_features = batch.read_level_features (base_classes.py:29-36).
 *   mdk_rl_create   lstm_size = cnn_size = 128 (the class defaults), 5 classes, kernel sizes [1, 17], mean pooling,
 *                   bidirectional; anything else -> MDK_ERR_UNSUPPORTED
 *   mdk_rl_load     one state-dict tensor by its torch name ("base_embedder.weight", "read_level_conv.convs.0.weight",
 *                   "read_level_conv.convs.2.running_mean", "lstm.weight_ih_l0_reverse", "linear.bias", ...), host float32,
 *                   torch's own layouts; tensors the forward does not use (num_batches_tracked,
 *                   read_level_conv.expansion_layer.*) may be skipped
 *   mdk_rl_forward  x int8 [B][P][D][F] host (the padded read-level feature tensor, torch_ext.py:127-136: F = 4, or 5 with
 *                   dwells) -> probs float32 [B][P][5] host (softmax output, normalise = True as at inference) */
typedef struct mdk_rl_engine mdk_rl_engine;
int mdk_rl_create(int device, int32_t lstm_size, int32_t cnn_size, int32_t use_dwells, int32_t num_classes,
                  mdk_rl_engine **out);
int mdk_rl_destroy(mdk_rl_engine *e);
int mdk_rl_load(mdk_rl_engine *e, const char *name, const float *data, int64_t n);
/* which parts run on tcgen05 with fp16 hi/lo operand pairs (bit set) or on the fp32 CUDA cores (validation twins):
 * bit 0 = the k = 17 convolution (99 % of the network's FLOPs), bit 1 = the LSTM recurrences.  Default 3. */
int mdk_rl_set_conv(mdk_rl_engine *e, int tensor_cores);
int mdk_rl_forward(mdk_rl_engine *e, const int8_t *x_host, int64_t B, int64_t P, int64_t D, int64_t F,
                   float *probs_host);

/* ---- alignment access: what calculate_pileup gets from htslib (create_bam_fset src/medaka_bamiter.c:52-63,
 * bam_itr_querys src/medaka_counts.c:233, the flag / mapQ part of read_bam src/medaka_bamiter.c:19-21).  Native BGZF
 * inflate (zlib, a thread pool over the independent members), BAI-indexed region fetch (bins + linear index; without an
 * index the file is streamed once in bounded memory), CG-tag long CIGARs resolved like htslib does.
 *   mdk_bam_open     index_path NULL = "<path>.bai" (or "<stem>.bai") when it exists
 *   mdk_bam_fetch    records of reference `tid` overlapping [start, end) whose flag has none of `exclude_flags` and
 *                    whose mapping quality is >= min_mapq, in file (coordinate) order
 *   mdk_bam_batch_arrays   pos[n], flag[n], mapq[n], l_seq[n]; cigar[] (len << 4 | op) with cigar_off[n+1]; seq[] 4-bit
 *                    codes with seq_off[n+1] (bytes); aux[] the raw optional fields with aux_off[n+1] (tag / read-group /
 *                    datatype filters are the caller's); names[] with name_off[n+1].  Any pointer may be NULL.  The
 *                    arrays live until mdk_bam_batch_free.
 * One handle may be used from several threads (reads are serialised on it). */
int mdk_bam_open(const char *path, const char *index_path, mdk_bam **out);
int mdk_bam_close(mdk_bam *b);
int mdk_bam_n_refs(mdk_bam *b);
const char *mdk_bam_ref_name(mdk_bam *b, int i);
int32_t mdk_bam_ref_len(mdk_bam *b, int i);
int mdk_bam_has_index(mdk_bam *b);
int mdk_bam_fetch(mdk_bam *b, int tid, int32_t start, int32_t end, uint32_t exclude_flags, int min_mapq, int threads,
                  mdk_bam_batch **out);
int64_t mdk_bam_batch_size(mdk_bam_batch *x);
int mdk_bam_batch_arrays(mdk_bam_batch *x, const int32_t **pos, const uint16_t **flag, const uint8_t **mapq,
                         const int32_t **l_seq, const uint32_t **cigar, const int64_t **cigar_off,
                         const uint8_t **seq, const int64_t **seq_off, const uint8_t **aux, const int64_t **aux_off,
                         const char **names, const int64_t **name_off);
/* base qualities as stored (bam_get_qual, src/medaka_read_matrix.c:428): l_seq bytes per read, 0xff when absent;
 * qual_off[n+1] */
int mdk_bam_batch_qual(mdk_bam_batch *x, const uint8_t **qual, const int64_t **qual_off);
int mdk_bam_batch_free(mdk_bam_batch *x);

/* ---- decode seam: replaces the array part of HaploidLabelScheme.decode_consensus ------------
 * (medaka/labels.py:1053-1085 with _phred :387-401): labels = argmax (first max wins),
 * quals = uint8(min(70, -10*log10(clip(1-p_max, 1e-7, 1)))) + 33.  probs float32 [n][5].
 * Gap removal / string building stays on the host.  quals may be NULL. */
int mdk_decode_consensus(int device, const float *probs, int64_t n, uint8_t *labels_out,
                         uint8_t *quals_out);
int mdk_decode_consensus_dev(int device, const float *probs_dev, int64_t n,
                             uint8_t *labels_out_dev, uint8_t *quals_out_dev);
/* same for float64 probabilities: all arithmetic in double, as numpy does for a float64 label_probs array */
int mdk_decode_consensus_f64(int device, const double *probs, int64_t n, uint8_t *labels_out,
                             uint8_t *quals_out);

/* Consensus stitching (medaka/stitch.py:33-85 `_stitch_samples`: per trimmed sample decode_consensus(with_qualities)
 * -> gap removal -> append).  The caller plans the kept row ranges (overlap trimming medaka/common.py:327-427,495-557,
 * region trimming :560-609, depth filter :612-644) and passes one pointer per range; the library copies only those
 * rows in, decodes them, removes gap calls and returns the concatenated ASCII bases / phred+33 quality characters.
 *   seg_probs[k]   host float32 [seg_rows[k]][5], first kept row of range k;  seg_rows[k] > 0
 *   seq_out/qual_out  host, capacity sum(seg_rows) bytes (qual_out may be NULL)
 *   seg_out_off    host [n_seg + 1]: range k produced seq_out[seg_out_off[k] : seg_out_off[k+1]]
 * The _dev form takes the ranges already back to back on the device (probs_dev [n_rows][5], 16-byte aligned) with
 * seg_base[k] = first row of range k (host array, strictly increasing from 0); seq/qual outputs are device pointers
 * of capacity n_rows, seg_out_off stays a host array. */
int mdk_stitch_consensus(int device, const float *const *seg_probs, const int64_t *seg_rows, int64_t n_seg,
                         uint8_t *seq_out, uint8_t *qual_out, int64_t *seg_out_off);
int mdk_stitch_consensus_dev(int device, const float *probs_dev, int64_t n_rows, const int64_t *seg_base,
                             int64_t n_seg, uint8_t *seq_out_dev, uint8_t *qual_out_dev, int64_t *seg_out_off);

/* variant_columns (src/medaka_rnn_variants.h:26, called at medaka/labels.py:869-887): which pileup columns belong
 * to a variant run.  minor [len] pileup minor indices; reference / prediction [len] one byte per column (the
 * symbol or label code incl. the gap - the reference passes wchar_t strings, any 1-byte coding with the same
 * equalities works); out [len] 0/1.  Host pointers. */
int mdk_variant_columns(int device, const int64_t *minor, const uint8_t *reference,
                        const uint8_t *prediction, uint8_t *out, int64_t len);

/* Variant decoding (medaka/labels.py:889-1014 `HaploidLabelScheme.decode_variants`, the array half): for one joined
 * sample of n pileup columns
 *   pred[i]      = argmax label of probs[i] (gaps kept; labels.py:917)
 *   is_var[i]    = variant_columns(minor, reference-with-gaps, prediction)   (src/medaka_rnn_variants.c:28-55)
 *   pred_q[i]    = phred(1 - probs[i][pred[i]]),  ref_q[i] = phred(1 - probs[i][ref_code[i]])   (labels.py:387-401,
 *                  float32; 'N' is scored as the gap class, labels.py:949-952)
 *   runs         = maximal runs of variant columns (common.rle, labels.py:928-930): first column, length and the
 *                  left-to-right float32 sums of pred_q / ref_q over the run (labels.py:957-975; the variant's
 *                  quality is run_pred_q - run_ref_q)
 * ref_code[i]: 0..4 = '*ACGT' with 0 on insertion columns (minor != 0), 5 = 'N', 6+ = any other draft symbol.
 * Strings, the ref == alt / ambiguous-reference filters and VCF normalisation stay on the host (medaka_b200/labels.py).
 * Host pointers; pred_q_out / ref_q_out may be NULL.  *n_runs_out is always set; if it exceeds max_runs the call
 * returns MDK_ERR_NOMEM without run data and the caller retries with larger buffers. */
int mdk_decode_variants(int device, const float *probs, const int64_t *minor, const uint8_t *ref_code, int64_t n,
                        uint8_t *pred_out, uint8_t *is_var_out, float *pred_q_out, float *ref_q_out, int64_t max_runs,
                        int64_t *run_start, int64_t *run_len, float *run_pred_q, float *run_ref_q,
                        int64_t *n_runs_out);

/* ---- self test of the tcgen05 building block (one 128xN tile GEMM), used by tests ----------
 * Computes D[128][N] = A[128][K] * B[N][K]^T with the same smem layouts, descriptors and
 * fp16 hi/lo split the GRU kernels use.  A, B, D are host fp32.  variant selects descriptor
 * hypotheses (0 = production encoding). */
int mdk_selftest_umma(int device, const float *A, const float *B, float *D, int N, int K,
                      int variant);

/* ---- diagnostics: cycle stamps of the recurrent kernel's hand-off points ------------------
 * enable != 0 switches the (slower, instrumented) recurrent kernels on for subsequent forwards on `device` at batch
 * sizes that run one tile per CTA; enable == 0 switches back.  If out is not NULL it first receives the stamps of
 * the last traced forward: uint64 [2 layers][16 time steps (512..527)][32 slots] of %clock64 on CTA (0,0); the slot
 * meanings are listed in tools/diag.py.  Not part of the hot path. */
int mdk_debug_rec_trace(int device, int enable, uint64_t *out);
/* diagnostics of the TRACED ping-pong recurrent kernels: a bit set that switches parts of a time step off (1: h-tile
 * stores, 2: proxy fence, 4: gate arithmetic, 8: x staging, 16: gi staging, 32: tile copy-out) so that the cycle trace
 * shows what each costs; results are wrong while any bit is set.  0 restores normal operation. */
int mdk_debug_pp_flags(int flags);
/* completion times (ms after mdk_engine_timer_start's event) of the eight stage events of the last n_last forwards,
 * oldest first: out is float[n_last][8] = start, features in, inproj0, rec0, inproj1, rec1, head, end.  Shows how the
 * lanes' kernels actually interleaved. */
int mdk_debug_timeline(mdk_engine *e, int n_last, float *out);
/* partial logits of the last forward on the fused-head path: float32 [2 directions][tiles][T][5 classes][16 windows]
 * (what the layer-1 recurrence writes instead of h1); per-direction parity checks of the fused linear head */
int mdk_debug_read_plog(mdk_engine *e, float *out_host, int64_t n_floats);

#ifdef __cplusplus
}
#endif
#endif /* MEDAKA_B200_H */
