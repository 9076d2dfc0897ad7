/* vcagan.h -- C ABI of libvcagan_b200.so (sm_100a).
 *
 * The reference (ms-dot-k/Visual-Context-Attentional-GAN) has no FFI layer: its hot path is the set of
 * torch.nn modules imported at train.py:7-8 / test.py:7-8 (src/models/{visual_front,resnet,generator}.py) plus
 * griffin_lim (src/data/audio_processing.py:51-68).  Each entry point below replaces the device work behind one
 * family of torch calls in those files (cited per group); the Python facade in
 * visual-context-attentional-gan_b200/src/models keeps the reference class names/signatures and calls these
 * through ctypes.
 *
 * Conventions: plain pointers (device memory unless noted) and sizes; no hidden allocation; no host sync;
 * every call enqueues on `stream`; returns 0 or a negative VCA_ERR_* code, message via vca_last_error().
 * Activations are channels-last: [N, D, H, W, C] contiguous (D = 1 for 2-D, D = H = 1 for 1-D / linear).
 * dtype codes: 0 = fp32, 1 = bf16 (fp32 accumulate everywhere).
 */
#ifndef VCAGAN_H_
#define VCAGAN_H_

#ifdef __cplusplus
extern "C" {
#endif

#ifndef __CUDA_RUNTIME_H__
typedef struct CUstream_st* cudaStream_t;
#endif

typedef struct ConvGeom {
  int N, ID, IH, IW, Cin;
  int OD, OH, OW, Cout;
  int KD, KH, KW;
  int sd, sh, sw;
  int pd, ph, pw;
} ConvGeom;

/* ---- library -------------------------------------------------------------------------------------------- */
const char* vca_last_error();
int vca_abi_version();
int vca_device_ok();
int vca_set_option(const char* key, int value);

/* ---- convolutions (nn.Conv1d/2d/3d: visual_front.py:11, resnet.py:5-14, generator.py:19-25,62-66,103-108,
 *      177-185,204-225,272-300,323-327) and their autograd (aten::convolution_backward) ------------------- */
int vca_pack_conv_weight(int dtype, const float* w, void* wf, void* wd, int Cout, int Cin, int taps, cudaStream_t stream);
/* every conv weight of an optimizer group in one launch: jobs = device table, 8 x int64 per job {w, wf, wd, Cout, Cin, taps,
   first CTA (prefix sum of vca_pack_job_ctas), 0}; total_ctas = the sum */
int vca_pack_job_ctas(int Cout, int Cin, int taps);
int vca_pack_conv_weights_batched(int dtype, const long long* jobs, int njobs, long long total_ctas, cudaStream_t stream);
/* Pixel-pair merge of a (Cin -> Cout, KH x 5, pad 2, stride 1) convolution (the 32-channel nn.Conv2d(.., 5, 1, 2) of
 * generator.py:58-60): the same linear map as a (2Cin -> 2Cout, KH x 3, pad 1) convolution over the free view
 * [N,H,W/2,2Cin], with w2[p*Cout+co][s*Cin+ci][kh][jj] = w[co][ci][kh][2jj+s-p].  expand: w -> w2 (fp32);
 * contract: the adjoint, dw2 -> dw (accumulate != 0 adds to dw). */
int vca_pair_expand_weight(const float* w, float* w2, int Cout, int Cin, int KH, cudaStream_t stream);
int vca_pair_contract_wgrad(const float* dw2, float* dw, int Cout, int Cin, int KH, int accumulate, cudaStream_t stream);
int vca_conv_fwd_simt(int dtype, const ConvGeom* g, const void* x, const void* wf, const float* bias, void* y, cudaStream_t stream);
int vca_conv_dgrad_simt(int dtype, const ConvGeom* g, const void* dy, const void* wd, void* dx, cudaStream_t stream);
int vca_conv_wgrad_simt(int dtype, const ConvGeom* g, const void* dy, const void* x, float* dw, cudaStream_t stream);
/* tcgen05/TMEM/TMA implicit GEMM, bf16 in / fp32 accumulate / bf16 out, stride 1 (conv_tc.cu) */
int vca_conv_tc_supported(const ConvGeom* g, int kind);
int vca_conv_fwd_tc(const ConvGeom* g, const void* x, const void* wd, const float* bias, void* y, cudaStream_t stream);
int vca_conv_dgrad_tc(const ConvGeom* g, const void* dy, const void* wf, void* dx, cudaStream_t stream);
/* Split-K variants for small-grid, long-K problems (the 5x5 heads of the discriminators on 5x18 maps): the caller lends
 * an fp32 workspace of vca_conv_tc_workspace(g, kind) bytes (0 = the plain entry points do the same work; kind 0 =
 * forward, 1 = dgrad); partial sums are red.add-ed into it and a finishing pass adds the bias and writes bf16. */
int vca_conv_tc_workspace(const ConvGeom* g, int kind);
int vca_conv_fwd_tc_ws(const ConvGeom* g, const void* x, const void* wd, const float* bias, void* y, float* ws, long long ws_bytes, cudaStream_t stream);
/* forward conv that also ADDS the per-output-channel sum / sum of squares of y (as stored) into stats[0..Cout) / stats[Cout..2Cout)
   (fp64): the batch statistics of a BatchNorm that follows (generator.py:115-116, resnet.py:47-48); never split-K */
/* inference forward with a fused epilogue: y = act(conv(x,w) * scale[c] + shift[c] + res * res_scale); act 0 none, 1 LeakyReLU(slope),
   2 PReLU(prelu_w[Cout]), 3 ReLU; scale / shift / res may be null.  Eval-mode BatchNorm folded into the conv (test.py:126-141). */
int vca_conv_fwd_tc_epi(const ConvGeom* g, const void* x, const void* wd, const float* scale, const float* shift, const void* res, float res_scale, int act, float slope, const float* prelu_w, void* y, cudaStream_t stream);
/* scale[c] = gamma[c] / sqrt(running_var[c] + eps); shift[c] = beta[c] + (bias[c] - running_mean[c]) * scale[c]  (bias may be null) */
int vca_bn_fold(const float* running_mean, const float* running_var, const float* gamma, const float* beta, const float* bias, int C, float eps, float* scale, float* shift, cudaStream_t stream);
/* inference: eval-mode BatchNorm folded into weights (scale) + bias (shift), y = a(conv + shift), a(v) = v > 0 ? v : v * slope[c]
 * (test.py:126-141 runs the modules in eval mode); weights-stationary geometries only, else VCA_ERR_UNSUPPORTED */
int vca_conv_fwd_tc_act_supported(const ConvGeom* g);
int vca_conv_fwd_tc_act(const ConvGeom* g, const void* x, const void* wd, const float* shift, const float* slope, void* y, cudaStream_t stream);
int vca_conv_fwd_tc_stats_supported(const ConvGeom* g);
int vca_conv_fwd_tc_stats(const ConvGeom* g, const void* x, const void* wd, const float* bias, void* y, double* stats, cudaStream_t stream);
int vca_conv_dgrad_tc_ws(const ConvGeom* g, const void* dy, const void* wf, void* dx, float* ws, long long ws_bytes, cudaStream_t stream);
int vca_conv_wgrad_tc(const ConvGeom* g, const void* dy, const void* x, float* dw, cudaStream_t stream);
/* the same weight gradient ADDED to a TAP-MAJOR fp32 tensor [taps][Cout][Cin] (16-byte aligned, Cin % 4 == 0) through shared
 * memory + TMA reduce-add (no scattered atomics); vca_grad_unslab_batched folds the slabs of an optimizer group back into
 * the parameter layout ([Cout][Cin][taps], what torch.autograd leaves in train.py:210,236's .grad) and zeroes them */
int vca_conv_wgrad_tc_tm(const ConvGeom* g, const void* dy, const void* x, float* dw_tm, cudaStream_t stream);
int vca_grad_unslab_batched(const long long* jobs, int njobs, long long total_ctas, int max_taps, cudaStream_t stream);
int vca_unslab_job_ctas(int Cout, int Cin, int taps);


/* ---- GEMM (nn.Linear: visual_front.py:21, generator.py:147-152,293,303,336; torch.bmm: generator.py:161,167,354;
 *      nn.GRU projections: visual_front.py:20) ---------------------------------------------------------------- */
int vca_gemm_simt(int dtA, int dtB, int dtC, const void* A, const void* B, void* C, const float* bias, int Z, int M, int N, int K, long long sAz, long long sAm, long long sAk, long long sBz, long long sBk, long long sBn, long long sCz, long long sCm, long long sCn, float alpha, float beta, cudaStream_t stream);

/* ---- BatchNorm (+residual) (+PReLU/LeakyReLU/ReLU), activations (visual_front.py:12-13, resnet.py:34-63,
 *      generator.py:9,52,95,105-126,179,209-225,325-329) --------------------------------------------------- */
/* sums_prezeroed / flags bit 0: `sums` is a persistent scratch that is zero on entry and left zeroed (no memset launch);
 * flags bit 1 of vca_bn_act_bwd: dgamma / dbeta / dprelu are added to in place (gradient accumulation). */
int vca_bn_stats(int dtype, const void* x, long long R, int C, float eps, float momentum, double* sums, int sums_prezeroed, float* mean, float* invstd, float* running_mean, float* running_var, cudaStream_t stream);
/* mean / invstd / running-buffer update from statistics a producer kernel accumulated; sums = double[2*C*fold] ([fold][C] sums,
   [fold][C] sums of squares), left zeroed */
int vca_bn_finalize_stats(double* sums, long long R, int C, int fold, float eps, float momentum, float* mean, float* invstd, float* running_mean, float* running_var, cudaStream_t stream);
int vca_bn_eval_stats(const float* running_mean, const float* running_var, int C, float eps, float* mean, float* invstd, cudaStream_t stream);
/* visual front-end stem tail in one pass: BatchNorm3d -> PReLU -> MaxPool3d (1,3,3)/(1,2,2)/(0,1,1) (visual_front.py:12-14), bf16.
   x [NF,H,W,C] raw conv output; y / idx / xmax [NF,OH,OW,C] (pooled activation, argmax code 0..8, raw x at the argmax). */
int vca_bn_prelu_maxpool_fwd(const void* x, void* y, unsigned char* idx, void* xmax, int NF, int H, int W, int C, const float* mean, const float* invstd, const float* gamma, const float* beta, const float* prelu_w, cudaStream_t stream);
int vca_bn_prelu_maxpool_bwd(const void* dy, const unsigned char* idx, const void* xmax, const void* x, void* dx, int NF, int H, int W, int C, const float* mean, const float* invstd, const float* gamma, const float* beta, const float* prelu_w, int train, double* sums, float* dgamma, float* dbeta, float* dprelu, int flags, cudaStream_t stream);
int vca_bn_act_fwd(int dtype, const void* x, const void* res, void* y, long long R, int C, const float* mean, const float* invstd, const float* gamma, const float* beta, int act, float slope, const float* prelu_w, cudaStream_t stream);
int vca_bn_act_bwd(int dtype, const void* dy, const void* x, const void* res, void* dx, void* dres, long long R, int C, const float* mean, const float* invstd, const float* gamma, const float* beta, int act, float slope, const float* prelu_w, int train, double* sums, float* dgamma, float* dbeta, float* dprelu, int flags, cudaStream_t stream);
int vca_lrelu_fwd(int dtype, const void* x, void* y, long long n, float slope, cudaStream_t stream);
int vca_lrelu_bwd(int dtype, const void* dy, const void* x, void* dx, long long n, float slope, cudaStream_t stream);
int vca_tanh_fwd(int dtype, const void* x, void* y, long long n, cudaStream_t stream);
int vca_tanh_bwd(int dtype, const void* dy, const void* y, void* dx, long long n, cudaStream_t stream);
int vca_axpby(int dtype, const void* a, const void* b, void* out, long long n, float alpha, float beta, cudaStream_t stream);
int vca_colsum(int dtype, const void* x, long long R, int C, double* scratch, float* out, int accumulate, cudaStream_t stream);
/* Row-tap combine of a KH x KW conv over channels that are constant along H (the phoneme features tiled along the mel
 * axis, generator.py:249-250): y[b,f,t,c] = yn[b,f,t,c] + sum_{kh: 0 <= f+kh-ph < F} R[b,t,kh*C+c];  bwd: dR = the
 * matching row sums of dy. */
int vca_row_taps_fwd(int dtype, const void* yn, const void* R, void* y, int B, int F, int T, int C, int KH, int ph, cudaStream_t stream);
int vca_row_taps_bwd(int dtype, const void* dy, void* dR, int B, int F, int T, int C, int KH, int ph, cudaStream_t stream);
int vca_cast(int dt_in, int dt_out, const void* x, void* y, long long n, cudaStream_t stream);
int vca_mul(int dtype, const void* x, const void* m, void* y, long long n, cudaStream_t stream);

/* ---- pooling / resampling (visual_front.py:14, resnet.py:82, generator.py:74,83,112,121,140) -------------- */
int vca_maxpool3x3s2_fwd(int dtype, const void* x, void* y, unsigned char* idx, int NF, int H, int W, int C, cudaStream_t stream);
int vca_maxpool3x3s2_bwd(int dtype, const void* dy, const unsigned char* idx, void* dx, int NF, int H, int W, int C, cudaStream_t stream);
int vca_pool2x2_sum(int dtype, const void* x, void* y, int NF, int H, int W, int C, float scale, cudaStream_t stream);
int vca_expand2x2(int dtype, const void* x, void* y, int NF, int IH, int IW, int C, int H, int W, float scale, cudaStream_t stream);
int vca_spatial_sum(int dtype, const void* x, void* y, int NF, int P, int C, float scale, cudaStream_t stream);
int vca_spatial_bcast(int dtype, const void* x, void* y, int NF, int P, int C, float scale, cudaStream_t stream);
/* layout transforms that put stride-2 convs (resnet.py:33, generator.py:323-327) and the Cin=1 stem (visual_front.py:11) on the tcgen05 path */
int vca_s2d(int dtype, const void* x, void* y, int NF, int H, int W, int C, int H2, int W2, cudaStream_t stream);
int vca_d2s(int dtype, const void* y, void* x, int NF, int H, int W, int C, int H2, int W2, cudaStream_t stream);
int vca_stem_im2col(int dt_in, int dt_out, const void* x, void* y, long long NF, int H, int W, cudaStream_t stream);

/* ---- batched bf16 GEMM on tcgen05 (attention backward contractions of generator.py:154-171, sync similarity :353) ----
 * C[z] (M x N row-major, ld ldc, batch stride strideC; bf16 or fp32) = alpha * op(A[z]) op(B[z]), fp32 accumulate.
 * a_mn = 0: A is [Z][M][K] (K contiguous, row pitch lda); a_mn = 1: A is [Z][K][M] (M contiguous).  Same for B with N.
 * lda / ldb / strideA / strideB in elements, multiples of 8. */
int vca_bmm_tc(const void* A, const void* B, void* C, int Z, int M, int N, int K, int a_mn, int b_mn, long long lda, long long strideA, long long ldb, long long strideB, long long ldc, long long strideC, int out_f32, float alpha, cudaStream_t stream);

/* ---- visual-context attention (generator.py:154-171): QK^T -> key mask -> softmax -> PV in one tcgen05 kernel ------------
 * Q [B][Tq][256], K, V [B][S][256] bf16 contiguous, S <= 256; lens int32 [B] (keys >= lens[b] are masked);
 * O [B][Tq][256] bf16; P [B][Tq][SP] bf16, SP = S rounded up to 16 (the softmax, saved for backward). */
int vca_att_tc_supported(int S, int d);
int vca_att_fwd_tc(const void* Q, const void* K, const void* V, const int* lens, void* O, void* P, int B, int Tq, int S, float scale, cudaStream_t stream);
/* dS = scale * P o (dP - rowsum(dP o P)); P bf16, dP fp32, dS bf16, all [rows][SP]; columns >= S written as 0 */
int vca_att_softmax_bwd(const void* P, const float* dP, void* dS, long long rows, int S, int SP, float scale, cudaStream_t stream);

/* ---- GRU gates (visual_front.py:20,33-34), attention softmax (generator.py:161-164), sync losses
 *      (generator.py:347-359), gan_loss (generator.py:363-366), L1 (train.py:226-229), Adam (train.py:82-83) - */
int vca_gru_gate_fwd(const float* gi, const float* gh, const float* bhh, const float* hprev, float* hnext, float* out, float* gates, int ndir, int T, int B, int H, int step, cudaStream_t stream);
int vca_gru_gate_bwd(const float* dout, float* dh_carry, const float* gates, const float* out, float* dgi, float* dgh, float* dgh_cur, int ndir, int T, int B, int H, int step, cudaStream_t stream);
int vca_gru_seq_fwd(const float* gi, const float* whh, const float* bhh, float* hbuf, float* out, float* gates, unsigned* bar, int ndir, int T, int B, int H, cudaStream_t stream);
int vca_gru_seq_bwd(const float* dout, const float* whh, const float* gates, const float* out, float* dgi, float* dgh, float* dhc, float* dghc, float* dhz, unsigned* bar, int ndir, int T, int B, int H, cudaStream_t stream);
/* Cluster plan of the GRU recurrence for batch B, hidden size H: info[6] = {CTAs per cluster (0 = cluster kernels not
 * applicable, vca_gru_seq_* then run as cooperative grids), batch rows per cluster, clusters a bidirectional layer needs,
 * co-resident clusters fwd, bwd, K slices of the forward matvec}. */
int vca_gru_cluster_query(int B, int H, int* info);
int vca_skinny_gemm(const float* in, const float* wt, float* out, int Z, int Bn, int N, int K, float beta, cudaStream_t stream);
int vca_masked_softmax_fwd(const float* x, float* p, const int* lens, int Z, int R, int S, cudaStream_t stream);
int vca_softmax_bwd(const float* dp, const float* p, float* dx, int rows, int S, cudaStream_t stream);
int vca_l2norm_fwd(const float* x, float* y, float* norms, int rows, int D, float eps, cudaStream_t stream);
int vca_l2norm_bwd(const float* dy, const float* y, const float* norms, float* dx, int rows, int D, float eps, cudaStream_t stream);
int vca_nce_diag(const float* sim, float* loss, float* dsim, int Bn, int S, cudaStream_t stream);
int vca_cos_abs_mean_fwd(const float* v, const float* a, float* loss, float* saved, int Bn, int S, int D, cudaStream_t stream);
int vca_cos_abs_mean_bwd(const float* dloss, const float* v, const float* a, const float* saved, float* da, float* dv, int Bn, int S, int D, cudaStream_t stream);
int vca_softplus_mean(const float* x, float* out, float* dx, int n, float sign, cudaStream_t stream);
int vca_reduce_l1_sq(int dtype, const void* a, const void* b, long long n, float scale, int mode, float* out, cudaStream_t stream);
int vca_l1_bwd(int dtype, const void* a, const void* b, const float* g, long long n, float scale, void* da, cudaStream_t stream);
int vca_adam_step(float* p, const float* g, float* m, float* v, float* vmax, long long n, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float gscale, cudaStream_t stream);
int vca_rng(int dtype, void* out, long long n, unsigned long long seed, unsigned long long offset, int mode, float param, cudaStream_t stream);
/* CUDA-graph-replayable variants: step counter, learning rate / RNG stream position live in device memory */
int vca_adam_step_dev(float* p, const float* g, float* m, float* v, float* vmax, long long n, const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, int* step_dev, float gscale, int bump, cudaStream_t stream);
int vca_rng_dev(int dtype, void* out, long long n, unsigned long long seed, unsigned long long* ctr_dev, int mode, float param, cudaStream_t stream);

/* ---- Griffin-Lim STFT / ISTFT (src/data/stft.py:70-129, src/data/audio_processing.py:51-68) ---------------------- */
int vca_gl_frames(int mode, const float* sig, const float* angles_t, const float* mag_t, float* frames, float* spec_out, int B, int T, int L, cudaStream_t stream);
int vca_gl_ola(const float* frames, float* sig_out, int B, int T, int L, cudaStream_t stream);
/* one whole Griffin-Lim iteration (audio_processing.py:62-66: transform -> angles -> inverse) in one kernel: sig_in normalised
 * (in_norm != 0) or the un-normalised overlap-add sums of a previous call; acc_out (ZERO on entry) receives the new sums;
 * vca_gl_normalize divides by the window envelope (stft.py:110-127) */
int vca_gl_iter(const float* sig_in, int in_norm, const float* mag_p, float* acc_out, int B, int T, int L, cudaStream_t stream);
int vca_gl_normalize(const float* acc, float* sig_out, int B, int T, int L, cudaStream_t stream);
/* out[row][k1*32 + lane] = in[row][k1 + 10*bitrev5(lane)], bin 320 last: the bin order vca_gl_frames reads with unit stride when (mode & 2) */
int vca_gl_permute_bins(const float* in, float* out, long long rows, cudaStream_t stream);

/* ---- waveform tail / mel front (src/data/vid_aud_grid.py:190-232, 291-307; src/data/vid_aud_lrs2.py:257-263) -------- */
/* out[b][f][t] = post(sum_k pre(in[b][k][t]) * w[k][f]);  pre 1: exp(in*pre_mul+pre_add);  post 0: *post_arg, 1: log(max(.,post_arg)) */
int vca_filterbank_apply(const float* in, const float* w, float* out, int B, int K, int F, int T, int pre, int post, float pre_mul, float pre_add, float post_arg, cudaStream_t stream);
int vca_exp_affine(const float* x, float* y, long long n, float mul, float add, float scale, cudaStream_t stream);
/* y[n] = x[n] + coef*y[n-1] in fp64 (scipy.signal.lfilter([1],[1,-coef])), clamped to [lo,hi] (np.clip) */
int vca_deemphasis_clip(const float* x, float* y, int B, int L, double coef, float lo, float hi, cudaStream_t stream);

/* ---- clip preprocessing of the loader (src/data/vid_aud_grid.py:94-121, src/data/vid_aud_lrs2.py:87-120) -------------- */
/* frames u8 [n][H][W][3] -> out f32 [n][OH][OW]: PIL crop + bilinear resize (fixed-point tables kx/bx, ky/by from the host) +
   hflip + luma + ToTensor + Normalize + erase box; meta int [n][10] = l,u,r,b, flip, ex0,ey0,ex1,ey1, valid.  Bit-exact. */
int vca_clip_preprocess(const unsigned char* frames, int n_frames, int H, int W, const int* meta, const int* kx, const int* bx, const int* ky, const int* by, int ksx, int ksy, int crop_w, int crop_h, int OW, int OH, float mean, float stdv, float* out, cudaStream_t stream);

/* ---- data-parallel gradient exchange (the replacement of nn.DataParallel's gather / re-broadcast, train.py:112-119): NCCL sum
 *      all-reduce of flat gradient buckets, one communicator per device of the calling process (one process per GPU).  NCCL is
 *      bound at run time from the libnccl.so.2 already in the process; vca_comm_available() == 0 when there is none. ---------- */
int vca_comm_available();
int vca_comm_unique_id(void* id128);                                   /* rank 0: the 128-byte rendezvous token */
int vca_comm_init(const void* unique_id, int rank, int world);         /* collective; binds to the current CUDA device */
int vca_allreduce_bucket(void* ptr, long long count, int dtype, cudaStream_t stream);   /* in place, SUM; VCA_F32 / VCA_BF16 */
int vca_comm_world();
int vca_comm_destroy();

#ifdef __cplusplus
}
#endif
#endif /* VCAGAN_H_ */
