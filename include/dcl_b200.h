/* dcl_b200.h - C ABI of the B200-native doubly contrastive loss library (libdcl_b200.so).
 *
 * The reference (andyj1/doubly-contrastive-semseg) has no FFI / operator layer: its boundary for
 * this path is two Python nn.Modules, `PixelContrastLoss` (utils/loss.py:250-415) and `SupConLoss`
 * (utils/loss.py:84-205), called from trainer.py:116-198.  The functions below are what those two
 * modules' forward/backward bind through ctypes; each comment names the
 * reference lines the call replaces.
 *
 * Conventions
 *   - every pointer is a caller-allocated DEVICE buffer unless it says "host";
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*), allocates nothing,
 *     never synchronises, and returns 0 on success or a negative dcl_status / positive
 *     cudaError_t; `dcl_last_error()` returns a thread-local message;
 *   - "anchor rows" are the rows of the N x N contrast matrix; they are stored as "F-tiles":
 *     128 rows x 128 channels of bf16 = 32 KiB each, in the 128-byte-swizzled image tcgen05.mma
 *     consumes directly (layout in csrc/dcl_ptx.cuh).  Rows are padded to a multiple of 128 with
 *     zero features and label -1; padded rows never contribute.
 *   - D (channels) is fixed at 128 = SwiftNet/ResNet-18 decoder width (`dim_in`, loss.py:98-101).
 */
#ifndef DCL_B200_H
#define DCL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCL_DIM 128
#define DCL_TILE_ROWS 128
#define DCL_TILE_BYTES 32768
#define DCL_MODE_PIXEL 0  /* lp_ij = l_ij - log(exp(l_ij) + sum_neg exp(l_ik))   loss.py:376-381 */
#define DCL_MODE_SUPCON 1 /* lp_ij = l_ij - log(sum_{k!=i} exp(l_ik))            loss.py:196-197 */
#define DCL_CHUNK_PIXELS 2048 /* pixels per sampler chunk (one CTA) */
#define DCL_HIST_BINS 512     /* 256 label values x {hard, easy} */

enum dcl_status {
    DCL_OK = 0,
    DCL_ERR_ARG = -1,       /* bad argument (shape, alignment, null) */
    DCL_ERR_WORKSPACE = -2, /* workspace too small */
    DCL_ERR_ARCH = -3,      /* device is not sm_100 */
    DCL_ERR_COMM = -4,      /* NCCL missing or failed (sharded step only) */
    DCL_ERR_LABEL = -5      /* a label outside 0..255 (the sampler codes labels in one byte) */
};

int dcl_version(void);
const char* dcl_last_error(void);
/* 0 iff the current device can run the kernels (compute capability 10.x). */
int dcl_check_device(void);
/* Diagnostics only: profiling switches for the contrast kernels (1 = skip epilogue math, 2 = skip
 * the S = F_I F_J^T MMAs, 4 = skip the dF MMAs: results are invalid while any of these is set;
 * 8 = run the pixel term's forward through the legacy three-sweep path, 16 = launch every kernel
 * plainly instead of as a programmatic dependent of the one before it; results stay valid for both).
 * Returns the previous value. */
int dcl_debug_flags(int flags);
/* Diagnostics only: device buffer (5*32*8 int64, or NULL to disable) that CTA 0 of the pipelined
 * kernels fills with per-role clock64 stamps of its first 32 tiles. */
int dcl_debug_trace(void* device_buffer);
/* Diagnostics only: device buffer of 4*256*4 int64; every CTA of the sweep-P, backward and k_rows kernels
 * records (globaltimer ns, clock64) at entry and exit.  NULL switches it off. */
int dcl_debug_cta_times(void* device_buffer);
/* Diagnostics / tests (host only, no device needed): the tile partition a sweep (backward == 0: units = pairs of row
 * blocks) or the backward (units = row blocks) uses for nI local row blocks, nJ column blocks and `ctas` CTAs.
 * begin [ctas+1]: first flat tile index (unit*nJ + k) of every CTA (begin[G] = units*nJ); unit_first / unit_nseg
 * [units]: first CTA and number of CTAs of every unit; meta [4]: G, exclusive mode, maxseg, units. */
int dcl_debug_partition(int nI, int nJ, int ctas, int backward, long long* begin, int* unit_first,
                        int* unit_nseg, int* meta);
/* Number of kernels one dcl_contrast_fwd (backward == 0) or dcl_contrast_bwd call launches for `mode`. */
int dcl_contrast_launches(int mode, int backward);

/* ---------------------------------------------------------------- sampler front end
 * Replaces loss.py:396-408 (argmax over classes, nearest down-sampling of labels) and the
 * counting half of _hard_anchor_sampling (loss.py:278-285, 308-312).
 *   labels   [B,H,W] int64      predict [B,C_cls,h,w] f32
 *   code     [B,h*w] u16 out :  low byte = down-sampled label (0..255), bit 8 = easy
 *                               (label == argmax), 0xFFFF = label outside 0..255
 *   chunk_hist [B,512,n_chunks] i32 out : per (image, bin = label*2+easy, chunk) pixel counts
 *   counts   [B,512] i32 out  : per-image totals per bin  (hard = bin label*2, easy = +1)
 *   n_chunks = ceil(h*w / DCL_CHUNK_PIXELS)
 */
int dcl_sample_classify(const int64_t* labels, const float* predict, int B, int H, int W, int h,
                        int w, int C_cls, uint16_t* code, int32_t* chunk_hist, int32_t* counts,
                        void* stream);

/* Rank -> pixel selection: replaces the `nonzero()[perm[:k]]` indexing of loss.py:308-331.
 *   req [N,4] i32 : (image, label, easy(0/1), rank) - "the rank-th pixel, in raster order, of
 *                   that image whose code matches"; rank comes from the host's randperm draw.
 *                   image < 0 marks a padding row.
 *   pix [N] i32 out : flat pixel id image*h*w + p, or -1 for padding rows.
 *   rowof [B*hw] i32 out, optional (NULL to skip): the inverse map, rowof[pix[n]] = n and -1 everywhere else; with
 *                   it dcl_gather_tiles / dcl_scatter_grad work chunk-wise in pixel order instead of row by row.
 */
int dcl_sample_select(const uint16_t* code, const int32_t* chunk_hist, int B, int hw,
                      const int32_t* req, int N, int32_t* pix, int32_t* rowof, void* stream);

/* Gather anchor rows into F-tiles: replaces the NCHW->NHWC copy + per-class advanced-index gather
 * of loss.py:409-410, :333.
 *   feats [B,128,h*w] f32 (NCHW), pix [n_pad] (from dcl_sample_select; -1 => zero row)
 *   tiles [n_pad/128] F-tiles out, sqnorm [n_pad] f32 out (|bf16(f)|^2, the row shift)
 *   rowof: optional inverse map from dcl_sample_select (NULL: one warp per row)
 */
int dcl_gather_tiles(const float* feats, int B, int hw, const int32_t* pix, int n_pad,
                     void* tiles, float* sqnorm, const int32_t* rowof, void* stream);

/* Same, from a dense row-major matrix Z [n,128] f32 (image-level term, loss.py:161). Rows >= n
 * are zero padding. */
int dcl_pack_rows(const float* Z, int n, int n_pad, void* tiles, float* sqnorm, void* stream);

/* Host-side (no GPU work): replay of torch's CPU mt19937 for the sampler's randperm draws,
 * loss.py:327-330.  `torch_rng_state` is the HOST buffer returned by torch.get_rng_state() (updated in
 * place; write it back with torch.set_rng_state).  For anchor a (reference order) writes
 * randperm(num_hard[a])[:keep_hard[a]] then randperm(num_easy[a])[:n_view-keep_hard[a]] into
 * ranks[a][0..n_view) and advances the generator exactly as the reference's calls would. */
int dcl_host_sample_ranks(void* torch_rng_state, size_t state_bytes, int A, int n_view,
                          const int64_t* num_hard, const int64_t* num_easy, const int64_t* keep_hard,
                          int64_t* ranks);

/* Host-side (no GPU work): the whole host half of _hard_anchor_sampling in one call - class list and
 * n_view (loss.py:278-291), split rule (:314-325), the randperm draws (:327-330, generator advanced
 * exactly as the reference's calls would) and the class-sorted device row layout that
 * dcl_sample_select / dcl_gather_tiles consume.
 *   counts [B][256][2] i32 host (from dcl_sample_classify), torch_rng_state as above
 *   info [4] out: A (anchors = reference `total_classes`), n_view, n (valid rows), n_pad
 *   image, cls, num_hard, num_easy, keep_hard [>= B*256] i64 out; ranks [>= max_samples] i64 out
 *   req [4*n_cap] i32 out (image, label, easy, rank; -1 = padding row), y [n_cap] i32 out,
 *   ref_row [n_cap] i64 out (row index v*A + a in the reference's ordering), anchor [n_cap] i64 out,
 *   n_cap >= max_samples rounded up to 128 (and >= 128)
 * Returns 0; 1 when no class qualifies (reference: `return None, None`, loss.py:287-288); 2 when the
 * split rule reaches the reference's "this shoud be never touched" branch (info = num_hard, num_easy,
 * n_view of the offending anchor); negative dcl_status on bad arguments. */
int dcl_host_plan_rows(const int32_t* counts, int B, int ignore_label, int max_samples, int max_views,
                       void* torch_rng_state, size_t state_bytes, int32_t* info, int64_t* image,
                       int64_t* cls, int64_t* num_hard, int64_t* num_easy, int64_t* keep_hard,
                       int64_t* ranks, int32_t* req, int32_t* y, int64_t* ref_row, int64_t* anchor);

/* Sharded form of dcl_host_plan_rows for one process per GPU (the reference has no multi-GPU path, SURVEY
 * D7): `counts` is the all-gathered table [world*Bl][256][2] in rank-major image order; every rank
 * replays the same generator stream but only draws the permutations of its own anchors.
 * info [6]: A, n_view, n (valid LOCAL rows), n_pad (common block size), n_global, -.  y_all
 * [world*n_pad] receives the labels of every rank's row block; req / ref_row / anchor [n_pad] describe
 * the local block (image index relative to the rank's first image).  y_all == NULL selects the compact
 * form: the labels are written right behind the requests, at req + 4*n_pad (req must then hold
 * (4 + world) * n_cap ints), so one host-to-device copy moves both. */
int dcl_host_plan_rows_sharded(const int32_t* counts, int Bl, int world, int rank, int ignore_label,
                               int max_samples, int max_views, void* torch_rng_state, size_t state_bytes,
                               int32_t* info, int64_t* image, int64_t* cls, int64_t* num_hard,
                               int64_t* num_easy, int64_t* keep_hard, int64_t* ranks, int32_t* req,
                               int32_t* y_all, int64_t* ref_row, int64_t* anchor);

/* Diagnostics: counters of the host generator look-ahead used by the two plan calls above (a worker
 * thread regenerates mt19937 state blocks ahead of the step while torch's generator stays where the
 * previous plan left it): out[4] = plans served by the stream, inline plans, stream starts, drops.
 * DCL_HOST_LOOKAHEAD=0 in the environment disables the stream. */
int dcl_host_lookahead_stats(long long* out);
/* Diagnostics: out[2] = nanoseconds the plans have spent waiting for the look-ahead worker (total, longest wait). */
int dcl_host_lookahead_wait(long long* out);
/* Diagnostics: cumulative nanoseconds of the last plan at the end of each of its sections: out[8] = anchor list,
 * + generator state / look-ahead attach, + permutations, + state write-back / commit, + row requests, -, -, -. */
int dcl_host_plan_timing(long long* out);

/* Diagnostics / tests (host only, no device needed): the host plan exactly as dcl_step_fwd requests it, with the
 * device-plan descriptors.  meta [8]: 1 if the permutations were left to the GPU, local anchors, stream epoch, first
 * and last stream block the permutations read, ring blocks, blocks produced so far, -.  `anchors`: 48-byte records
 * (u64 g_hard, u64 g_easy, i32 num_hard, num_easy, keep_hard, keep_easy, row0, image, cls, -): g_* = position of the
 * permutation's first draw in the generator's output stream, counted from word 0 of stream block 0.  ring_out: the
 * look-ahead ring of raw (untempered) mt19937 state blocks, block b at ring + (b % ring_blocks) * 624 u32 words.
 * ycls / yanchor: class and anchor id of the o-th class-sorted anchor of rank r at [yoff[r] + o], o < ycnt[r]. */
int dcl_debug_plan_device(const int32_t* counts, int Bl, int world, int rank, int ignore_label, int max_samples,
                          int max_views, void* torch_rng_state, size_t state_bytes, int32_t* info, int64_t* image,
                          int64_t* cls, int64_t* num_hard, int64_t* num_easy, int64_t* keep_hard, int64_t* ranks,
                          int32_t* req, int32_t* y_all, void* anchors, int32_t* ycls, int32_t* yanchor,
                          int32_t* ycnt, int32_t* yoff, long long* meta, const void** ring_out);

/* ---------------------------------------------------------------- N x N contrast
 * Forward of _contrastive (loss.py:339-389) / SupConLoss.forward (loss.py:175-204) for the local
 * row blocks [rb0, rb0+nI) against ALL nJ column blocks, N x N never materialised.
 *   tiles  [nJ] F-tiles (the whole contrast set; all-gathered by the caller when sharded)
 *   y      [nJ*128] i32 labels, -1 = padding        sqnorm [nJ*128] (read for the local rows only)
 *   n_valid : number of valid rows over the WHOLE contrast set (the reference's N)
 *   colA, colB [nJ*128] float4 out (rows of the local blocks only are written): per-row
 *            constants consumed by dcl_contrast_bwd - (a, b, p, q) and (wn, Den, label bits, logit range L);
 *            all-gather them before a sharded backward
 *   rowloss [nJ*128] f32 out (local rows): per-row loss term, 0 for padding
 *   loss_sum [2] f32 out: [0] sum of rowloss over the local rows, [1] that sum / n_valid (the loss when
 *            the rows are not sharded)
 *   workspace: dcl_contrast_workspace_bytes(nI, nJ) bytes
 */
size_t dcl_contrast_workspace_bytes(int nI, int nJ);
int dcl_contrast_fwd(const void* tiles, const int32_t* y, const float* sqnorm, int nJ, int rb0,
                     int nI, int n_valid, int mode, float temperature, float base_temperature,
                     void* workspace, size_t workspace_bytes, float* colA, float* colB,
                     float* rowloss, float* loss_sum, void* stream);

/* Backward: dF[local rows] = sum_k (dS_ik + dS_ki) F_k, tiles recomputed (autograd of
 * loss.py:361-388).  colA/colB must hold ALL rows.  dF [nI*128,128] f32 out (d loss / d row,
 * for loss = mean over n_valid rows; not yet multiplied by the upstream gradient). */
int dcl_contrast_bwd(const void* tiles, const int32_t* y, const float* colA, const float* colB,
                     int nJ, int rb0, int nI, int mode, void* workspace, size_t workspace_bytes,
                     float* dF, void* stream);

/* The same contrast for a handful of rows (n <= dcl_contrast_small_max_rows() = 128) in exact fp32, forward and
 * gradient in one launch: what the image-level term uses (2B rows; its rows are projections of pooled features,
 * nearly identical across images, and bf16 operands would drown their differences).  Z [n,128] f32 row-major,
 * y [n] i32 -> loss [1]; dZ [n,128] out (d loss / d Z, optional). */
int dcl_contrast_small_max_rows(void);
int dcl_contrast_small(const float* Z, const int32_t* y, int n, int mode, float temperature, float base_temperature,
                       float* loss, float* dZ, void* stream);

/* The rest of the image-level head at that size, so that the term is four launches instead of ~45 torch ones:
 * SupConLoss.projection (Linear 128->128, ReLU, Linear 128->128; reference loss.py:102-106, applied at :120) and the
 * group ids of the label mask (loss.py:151-159: mask = eq(labels, labels^T), or the identity without labels).
 *   X [n,128] f32 pooled rows; W1, W2 [128,128] f32 (nn.Linear weight, [out,in]); b1, b2 [128]
 *   H [n,128] out = relu(X W1^T + b1);  Z [n,128] out = H W2^T + b2
 *   backward: dZ [n,128] (unscaled, from dcl_contrast_small), grad_out [1] device scalar; dH [n,128] scratch;
 *             dX [n,128], dW1, dW2 [128,128], db1, db2 [128] out (all scaled by *grad_out)
 *   dcl_group_ids: labels [n] int32 / int64 (elem_bytes 4 / 8) or NULL; y [views*n] i32 out, y[v*n+i] = first k with
 *             labels[k] == labels[i] (i itself without labels) */
int dcl_supcon_mlp_fwd(const float* X, const float* W1, const float* b1, const float* W2, const float* b2, int n,
                       float* H, float* Z, void* stream);
int dcl_supcon_mlp_bwd(const float* X, const float* W1, const float* W2, const float* H, const float* dZ,
                       const float* grad_out, int n, float* dH, float* dX, float* dW1, float* db1, float* dW2,
                       float* db2, void* stream);
int dcl_group_ids(const void* labels, int elem_bytes, int n, int views, int32_t* y, void* stream);

/* ---------------------------------------------------------------- gradient back to NCHW
 * Replaces autograd of the gather (index_put into a zero tensor per class):
 * row n's gradient * (*grad_out) is written at pixel pix[n] of dfeats [B,128,h*w] f32 (pix < 0 skipped).
 * zero_fill: 0 = dfeats is already clear, 1 = clear it first, 2 = ADD to what dfeats holds (sampled pixels are
 * distinct, so no atomics are needed).  grad_out: device scalar (upstream dL/dloss).  rowof: optional inverse map
 * from dcl_sample_select (NULL: one warp per row). */
int dcl_scatter_grad(const float* dF, const int32_t* pix, int n_rows, const float* grad_out,
                     float* dfeats, int B, int hw, int zero_fill, const int32_t* rowof, void* stream);
/* Clears `bytes` (multiple of 16) at dst.  persistent != 0: two 128-thread blocks per SM walk the whole buffer, so
 * that on a second stream the fill never keeps kernels of the main stream waiting for a slot (the step clears the
 * dense gradient buffer this way while the count table is with the host); 0: one short block per 64 KB. */
int dcl_zero_fill(void* dst, size_t bytes, int persistent, void* stream);
/* The dense gradient of the doubly contrastive step (both losses hit the same `fine_feat`,
 * trainer.py:144-152): dfeats [B_all,128,hw] = gap_g[image*128 + channel] / hw everywhere (AdaptiveAvgPool2d
 * backward, loss.py:115), written once by the streaming broadcast, + (*grad_out) * dF[row] added at the sampled
 * pixels of the first B_pix images (rowof from dcl_sample_select).  No zero-fill; DRAM traffic 1.03x the tensor. */
int dcl_dense_grad(const float* dF, const int32_t* rowof, int B_pix, const float* grad_out, const float* gap_g,
                   float* dfeats, int B_all, int hw, void* stream);
/* dZ [n,128] = dF[:n] * (*grad_out)  (image-level term). */
int dcl_unpack_rows(const float* dF, int n, const float* grad_out, float* dZ, void* stream);

/* ---------------------------------------------------------------- sharded exchange (one process per GPU)
 * The one message a rank contributes before the backward: send [(2*n_pad + 1) float4] = colA of its rows | colB of
 * its rows | (its loss sum, 0, 0, 0).  After an all-gather of these messages, dcl_shard_unpack puts every rank's
 * constants in place (colA / colB [world*n_pad] float4) and writes loss[0] = sum of the partials (rank order) /
 * n_global.  New design: the reference has no multi-GPU path. */
int dcl_shard_pack(const float* colA, const float* colB, const float* loss_sum, int rank, int n_pad,
                   float* send, void* stream);
int dcl_shard_unpack(const float* recv, int world, int n_pad, float* colA, float* colB, int n_global,
                     float* loss, void* stream);

/* ---------------------------------------------------------------- global average pool
 * nn.AdaptiveAvgPool2d((1,1)) forward/backward (loss.py:104, :115): x [R,hw] rows = (image,
 * channel) pairs -> pooled [R]; backward broadcasts g[R]/hw. `accumulate` != 0 adds into dx
 * (fused with the pixel-term scatter target). */
int dcl_gap_fwd(const float* x, int R, int hw, float* pooled, void* stream);
int dcl_gap_bwd(const float* g, int R, int hw, float* dx, int accumulate, void* stream);

/* ---------------------------------------------------------------- the pixel term in one call
 * PixelContrastLoss.forward (loss.py:391-415 -> 250-389) AND the gradient of its N x N part, issued from C with one
 * host wait (the count table), for a single GPU (world == 1) or one rank of a row-sharded job (one process per GPU;
 * new design, the reference has no multi-GPU path):
 *   dcl_sample_classify -> [NCCL all-gather of the count tables] -> count table to the host
 *   (side stream: zero-fill of the dense gradient buffer) -> host: anchor list, n_view, split rule ->
 *   permutations on the GPU from a mirror of torch's CPU generator stream (k_plan; exact) or on the host when the
 *   generator was touched since the last step -> dcl_sample_select -> dcl_gather_tiles ->
 *   [NCCL all-gather of the F-tiles] -> dcl_contrast_fwd -> [NCCL all-gather of the row constants] ->
 *   dcl_contrast_bwd (eager: dF for an upstream gradient of 1; pass dF = NULL for a forward-only step).
 * The autograd backward is dcl_step_bwd: scale and scatter dF.  All buffers are the caller's:
 *   B                images of THIS rank; every rank passes the same B
 *   cap              rows per rank block the buffers hold: multiple of 128, >= max_samples rounded up to 128
 *   device scratch   code [B*h*w] u16, chunk_hist [B*n_chunks*512] i32, counts_dev [world*B*512] i32,
 *                    req_dev [4*cap] i32, y_dev [world*cap] i32, pix [cap] i32, rowof [B*h*w] i32 (pixel -> row
 *                    map; kept for dcl_step_bwd), plan_dev [plan_bytes]
 *   device state     tiles [world*cap*256 B], sqnorm [world*cap], colA/colB [world*cap*4] f32, rowloss [world*cap],
 *                    loss_sum [2], loss [1] (the step's result), dF [cap*128] f32 or NULL,
 *                    xchg_send [(2*cap+1)*4] f32 and xchg_recv [world*(2*cap+1)*4] f32 (world > 1 only)
 *   workspace        >= dcl_contrast_workspace_bytes(n, world*n) for every n <= cap/128
 *   pinned host      counts_host [world*B*512] i32, plan_host [plan_bytes = dcl_step_plan_bytes(B, world)],
 *                    stage_host [4*cap + world*cap] i32 (rows of a host-side plan)
 *   host outputs     info [8]: A (anchors = reference `total_classes`), n_view, n (valid local rows), n_pad (rows per
 *                    rank block), n_global, 1 if the permutations ran on the GPU, -, -;
 *                    image / cls / num_hard / num_easy / keep_hard [>= world*B*256] i64: the anchors in reference
 *                    order; ranks [>= cap] i64: randperm prefixes (host-side plan only)
 *   comm             dcl_comm_init handle (world > 1)
 *   p2p              optional dcl_p2p_create handle (world > 1): the three exchanges then go through NVLink peer memory
 *                    (push kernels + flag waits, csrc/dcl_p2p.cu) instead of NCCL, and counts_dev / tiles / xchg_* are
 *                    not used
 *   zero_fill        optional device buffer cleared off the critical path (side_stream: optional second stream)
 *   device_plan      0: always replay the generator on the host
 *   ev_*             optional cudaEvent_t recorded around the N x N forward / backward (measurement)
 * Returns 0; 1 when no class qualifies (reference: `return None, None`, loss.py:287-288); 2 for the reference's
 * "this shoud be never touched" branch (info = num_hard, num_easy, n_view); 3 when max_samples / total_classes == 0;
 * negative dcl_status on errors. */
typedef struct dcl_step {
    const int64_t* labels; const float* predict; const float* feats;
    int B, H, W, h, w, C_cls, ignore_label, max_samples, max_views;
    float temperature, base_temperature;
    void* torch_rng_state; size_t state_bytes;
    int world, rank; void* comm; void* p2p;
    int cap;
    uint16_t* code; int32_t* chunk_hist; int32_t* counts_dev;
    int32_t* req_dev; int32_t* y_dev; int32_t* pix; int32_t* rowof; void* plan_dev;
    void* tiles; float* sqnorm; float* colA; float* colB; float* rowloss; float* loss_sum; float* loss;
    float* xchg_send; float* xchg_recv; float* dF;
    void* workspace; size_t workspace_bytes;
    int32_t* counts_host; void* plan_host; size_t plan_bytes; int32_t* stage_host;
    int32_t* info; int64_t* image; int64_t* cls; int64_t* num_hard; int64_t* num_easy; int64_t* keep_hard;
    int64_t* ranks;
    void* zero_fill; size_t zero_fill_bytes; void* side_stream;
    int device_plan;
    void* ev_fwd_begin; void* ev_fwd_end; void* ev_bwd_begin; void* ev_bwd_end;
    int begun;          /* != 0: dcl_step_begin was already issued for this step on the same stream and thread */
} dcl_step_t;
size_t dcl_step_plan_bytes(int B_local, int world);
/* Optional first half of dcl_step_fwd (classify, count tables, zero-fill): issue it as soon as the inputs are known,
 * prepare the rest of the descriptor while the GPU classifies, then call dcl_step_fwd with begun = 1. */
int dcl_step_begin(const dcl_step_t* step, void* stream);
int dcl_step_fwd(const dcl_step_t* step, void* stream);
/* Autograd backward of the step: dfeats [B(+),128,hw] = (*grad_out) * scatter(dF at pix).  zero_fill != 0 clears
 * dfeats first (0: the caller already did, e.g. through dcl_step_t.zero_fill).  gap_g != NULL fuses the image-level
 * term's gradient into the same pass (both losses hit the same `fine_feat`, trainer.py:144-152):
 * dfeats has gap_rows/128 >= B images, every (image, channel) row r is written with gap_g[r] / hw (the
 * AdaptiveAvgPool2d backward, loss.py:115) and the anchor gradients are added at the sampled pixels (dcl_dense_grad):
 * the dense tensor is written once, no zero-fill. */
int dcl_step_bwd(const float* dF, const int32_t* pix, const int32_t* rowof, int n_pad, const float* grad_out,
                 float* dfeats, int B, int hw, int zero_fill, const float* gap_g, int gap_rows, void* stream);

/* Diagnostics: cumulative host nanoseconds of the last dcl_step_fwd at the end of each of its sections: out[8] =
 * classify / count-table copy / zero-fill issued, + wait for the count table, + host plan, + generator blocks queued
 * for upload, + plan kernel (or row copies) issued, + select / gather / forward issued, + exchange / backward issued, -. */
int dcl_step_timing(long long* out);

/* Measurement aid (bench.py): similarity-kernel time of the steps.  With timing enabled, every dcl_step_fwd on the
 * current device whose descriptor carries no events of its own brackets its contrast forward and backward with events
 * owned by the library (a ring of 2048 steps).  dcl_step_sim_timing returns the previous setting and resets the ring
 * when enabling; dcl_step_sim_elapsed sums the recorded intervals (milliseconds) of the last min(steps, 2048) steps -
 * synchronize the stream first. */
int dcl_step_sim_timing(int enable);
int dcl_step_sim_elapsed(double* fwd_ms, double* bwd_ms, long long* steps);

/* Measurement aid (bench.py): a native thread samples the CURRENT device's SM clock, throttle reasons and power through
 * NVML every period_us microseconds while a timed region runs (`nvidia-smi -lms` itself slows a sub-millisecond step).
 * dcl_clock_sampler_stop: out[5] = samples, median SM MHz, max SM MHz, OR of NVML throttle-reason masks (0x4 sw power
 * cap, 0x8 hw slowdown, 0x20 sw thermal slowdown, 0x40 hw thermal slowdown), highest power draw in W. */
int dcl_clock_sampler_start(int period_us);
int dcl_clock_sampler_stop(double* out);

/* ---------------------------------------------------------------- NCCL plumbing of the sharded step
 * libnccl.so.2 is taken from the process (torch has it mapped) or dlopen'ed; nothing is linked.  Rank 0 creates the
 * 128-byte id, every rank receives it out of band (e.g. one torch.distributed broadcast) and calls dcl_comm_init on
 * its own device.  dcl_comm_all_gather: `bytes_per_rank` from send into recv [world*bytes_per_rank], in place when
 * send == recv + rank*bytes_per_rank. */
int dcl_comm_unique_id(void* out128);
int dcl_comm_init(const void* id128, int world, int rank, void** comm);
int dcl_comm_destroy(void* comm);
int dcl_comm_all_gather(void* comm, const void* send, void* recv, size_t bytes_per_rank, void* stream);

/* ---------------------------------------------------------------- peer-memory exchange of the sharded step
 * Every rank creates an arena sized for (B images per rank, `cap` rows per rank, `world` ranks <= 16) and publishes its
 * 64-byte CUDA IPC handle; after an out-of-band all-gather of the handles (world * 64 bytes, rank order) dcl_p2p_open
 * maps the peers' arenas.  All ranks must live on one node with peer access between their GPUs (NVLink / NVSwitch). */
int dcl_p2p_create(int world, int rank, int B_local, int cap, void** p2p_out, void* ipc_handle_out);
int dcl_p2p_open(void* p2p, const void* all_handles);
int dcl_p2p_destroy(void* p2p);

/* ---------------------------------------------------------------- segmentation-loss neighbour
 * BoundaryAwareFocalLoss (loss.py:27-80), forward and gradient in one pass, the up-sampled logits never formed:
 *   logits [B,C,h,w] f32 (C <= 32) pre-upsample (h == H, w == W allowed: full-resolution logits); target [B,H,W] i64,
 *   REWRITTEN in place like the reference does (ignore_id -> 0, loss.py:43); alpha [B,H,W] f32 = the batch's
 *   `label_distance_weight`; weight [C] class weights (modes 0 and 3).
 *   mode: 0 weight*alpha (default), 1 `plain_focal`, 2 `no_class_weights`, 3 `no_EDT` (loss.py:63-70)
 *   dlogits_unscaled [B,C,h,w] out: N * d loss / d logits;  loss_n [2] out: the loss (0 when N == 0) and
 *   N = #(alpha > 0);  workspace >= dcl_focal_workspace_bytes(B, h, w).
 * dcl_focal_bwd: dlogits [n] = dlogits_unscaled * (*grad_out) / N (zeros when N == 0). */
size_t dcl_focal_workspace_bytes(int B, int h, int w);
int dcl_focal_fwd(const float* logits, int64_t* target, const float* alpha, const float* weight, int B, int C, int h,
                  int w, int H, int W, int ignore_id, float gamma, int mode, float* dlogits_unscaled, float* loss_n,
                  void* workspace, size_t workspace_bytes, void* stream);
int dcl_focal_bwd(const float* dlogits_unscaled, const float* loss_n, const float* grad_out, float* dlogits, size_t n,
                  void* stream);

#ifdef __cplusplus
}
#endif
#endif /* DCL_B200_H */
