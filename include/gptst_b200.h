/* gptst_b200.h -- C ABI of the B200-native GPT-ST pre-training hot path (libgptst_b200.so).
 *
 * Every entry point:
 *   - takes raw DEVICE pointers to contiguous fp32 buffers (no torch types), plain int sizes and a
 *     cudaStream_t passed as void*;
 *   - is stream-ordered and asynchronous: no allocation, no host synchronisation, no global state, so the
 *     calls can be captured into a CUDA graph;
 *   - returns 0 on success, a positive cudaError_t on a CUDA failure, -1 for a NULL/empty argument and
 *     -2 for an unsupported shape (D must be 64 or 128, T == 12, H <= 15, prec in {1,3}).
 *
 * The reference (HKUDS/GPT-ST) has no FFI: the path is pure PyTorch.  Each function below replaces the
 * cited lines of /root/reference/model/Pretrain_model/GPTST.py; INTEGRATION.md shows the ctypes binding.
 *
 * Layout conventions (all row-major, last index fastest):
 *   activation  x, eb, out ...  (B, T, N, D)          slab = one (b,t) block of N x D
 *   c, dadj, dcr, ddadj         (B, T, H, N)          hyperedge-major incidence per slab
 *   s, v, dv, ds                (B, T, H, D)
 *   dyn, ddyn                   (B, HT, T*H)
 * prec: 1 = TF32 tensor-core operands, 3 = 3xTF32 split (fp32-faithful, default of the Python layer).
 */
#ifndef GPTST_B200_H
#define GPTST_B200_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- grouped D x D projection with per-group weights -------------------------------------------------
 * Y[g][r][:] = act( X[g][r][:] . W[g] + bias[g] (+ Res[g][r][:]) ),  element (g,r,j) at
 * base + g*group_stride + r*row_stride + j.  W (G,D,D) is [in][out]; bias (G,D) may be NULL; Res may be NULL.
 * act: 0 = identity, 1 = LeakyReLU(0.01).
 *   time-adaptive projection  GPTST.py:160-163 (hyperTem), :29-32 (MLP_RL):  G=B*T, R=N,   gs=N*D, rs=D
 *   node-adaptive projection  GPTST.py:137-141 (cap),      :24-27 (MLP_RL):  G=N,   R=B*T, gs=D,   rs=N*D   */
int gptst_gproj_fwd(const float* X, const float* W, const float* bias, const float* Res, float* Y, int G, int R,
                    long group_stride, long row_stride, int D, int act, int prec, void* stream);
/* Backward of the above.  dy = dY * act'(Y)  (Y may be NULL when act == 0).
 *   dX = dy . W^T ;  dW_part[s][g] = partial sum_r X^T dy ;  dbias_part[s][g] = partial sum_r dy ;  dRes = dy (may be NULL)
 * dW_part is (splits, G, D, D), dbias_part (splits, G, D); the caller sums over `splits`
 * (splits = gptst_gproj_splits(G, R, D) row-range CTAs per group keep the reduction deterministic).          */
int gptst_gproj_splits(int G, int R, int D);
int gptst_gproj_bwd(const float* dY, const float* Y, const float* X, const float* W, float* dX, float* dW_part,
                    float* dbias_part, float* dRes, int G, int R, long group_stride, long row_stride, int D, int act,
                    int prec, int splits, void* stream);

/* ---- temporal hypergraph two-hop of hyperTem, GPTST.py:156-158, as a per-node T x T mix ---------------
 * y[b,t,n,:] (+)= sum_t' M[n][t][t'] x[b,t',n,:]   (transpose != 0 uses M[n][t'][t]); M is (N,T,T).           */
int gptst_tmix(const float* x, const float* M, float* y, int B, int T, int N, int D, int transpose, int accumulate,
               void* stream);
/* dM_part[s][n][t][t'] = partial over batch split s of sum_{b,j} dy[b,t,n,j] x[b,t',n,j]                       */
int gptst_tmix_dM_splits(int B, int N);
int gptst_tmix_dM(const float* dy, const float* x, float* dM_part, int B, int T, int N, int D, int splits, void* stream);

/* backward of the mix, fused (D = 64): dx_io[b,s,n,:] += sum_t M[n][t][s] dy[b,t,n,:]  and  dM_part[p][n][t][s] = partial over the
 * batch range p of sum_{b,j} dy[b,t,n,j] x[b,s,n,j]; p < gptst_tmix_bwd_splits(B, N), summed by the caller.                 */
int gptst_tmix_bwd_splits(int B, int N);
int gptst_tmix_bwd(const float* dy, const float* x, const float* M, float* dx_io, float* dM_part, int B, int T, int N, int D,
                   int prec, int splits, void* stream);

/* ---- fused hyperTem block, GPTST.py:154-163 (D = 64, T = 12): one TMA-fed persistent kernel per direction --------------------
 * Replaces gptst_tmix + gptst_gproj_fwd (forward) and gptst_gproj_bwd + gptst_tmix_bwd (backward) on the main chain.
 *   pack_w : W (G = B*T, 64, 64) fp32 [in][out] -> fragment-ordered fp16 hi/lo tables, wf for the forward (ret W), wb for the
 *            backward (dy W^T); gptst_hypertem_wfrag_bytes(G) bytes each; either may be NULL.
 *   fwd    : out = LReLU((M_n o eb) W_bt + bias_bt + eb);  mask = sign words of out, (B*T, Npad, 2) uint32 with Npad = N rounded
 *            up to 16 and bit 16*(c & 3) + (c >> 2) of a row's 64 bits = out[c] > 0;  ret (may be NULL) = M_n o eb.
 *   bwd    : deb = dy + M_n^T o (dy W_bt^T) with dy = dOut . LReLU'(mask);  dret (may be NULL) = dy W_bt^T.
 *   dw     : dW_part[s][b,t] = partial of ret_bt^T dy_bt, dbias_part[s][b,t] = partial of sum_n dy_bt  (s < gptst_gproj_splits(B*T, N, 64))
 *   tmix_dM2: dM_part[s][n][t][t'] = partial of sum_{b,j} dret[b,t,n,j] eb[b,t',n,j]                  (s < gptst_tmix_bwd_splits(B, N))
 * The last two are the parameter-side gradients (G_bt, sigma_bt, dM_n); nothing on the main chain reads them. */
long gptst_hypertem_wfrag_bytes(int G);
int gptst_hypertem_mask_pad_rows(void);
int gptst_hypertem_pack_w(const float* W, void* wf, void* wb, int G, void* stream);
int gptst_hypertem_fwd(const float* eb, const float* Mn, const void* wfrag, const float* bias, float* out, void* mask, float* ret, int B,
                       int T, int N, int D, void* stream);
int gptst_hypertem_bwd(const float* dout, const void* mask, const float* Mn, const void* wfrag_t, float* deb, float* dret, int B, int T,
                       int N, int D, void* stream);
int gptst_hypertem_dw(const float* dout, const void* mask, const float* ret, float* dW_part, float* dbias_part, int B, int T, int N,
                      int D, int mask_rows, int splits, void* stream);
int gptst_tmix_dM2(const float* dret, const float* x, float* dM_part, int B, int T, int N, int D, int splits, void* stream);

/* ---- cap: intra-cluster routing, GPTST.py:102-123 ----------------------------------------------------
 * P = squash(x Wp^T + bp); R routing iterations on (P, dadj); c = softmax_H(b + dadj) -> c (B,T,H,N); s = c P.
 * D = 64, N <= 256: one CTA per (b,t) slab, one warp per 16 nodes, fp16-split tensor-core contractions
 * (cap_route2_fwd.cu).  Otherwise one thread-block cluster per slab, sized so the slab's P tile stays in shared memory
 * (cap_route_fwd.cu, tf32 split).                                                                               */
int gptst_cap_route_fwd(const float* x, const float* Wp, const float* bp, const float* dadj, float* c, float* s, int B,
                        int T, int N, int D, int H, int R, int prec, void* stream);
/* ---- cap: inter-cluster hop over k=(t,h) per sample, GPTST.py:125-134:  v = squash(LReLU(dyn^T LReLU(dyn (s+tau))) + s) */
int gptst_cap_hop_fwd(const float* s, const float* dyn, float* v, int B, int T, int D, int H, int HT, void* stream);
/* ---- cap: hyperedge -> node reconstruction, GPTST.py:135:  recon[b,t,n,:] = sum_h c[b,t,h,n] v[b,t,h,:]      */
int gptst_cap_recon(const float* c, const float* v, float* recon, int B, int T, int N, int D, int H, void* stream);
/* ---- cap: the same two steps (GPTST.py:125-135) split for parallelism; this pair is what the forward path launches.
 * hop_e1:    e1[b] = LReLU(dyn_b (s_b + tau)), e1 is (B, HT, D) -- the only part that mixes the T slabs of a sample.
 * recon_hop: per slab  v = squash(LReLU(dyn_b[:, t-block]^T e1[b]) + s) -> v (B,T,H,D);  recon = c^T v -> (B,T,N,D).   */
int gptst_cap_hop_e1(const float* s, const float* dyn, float* e1, int B, int T, int D, int H, int HT, void* stream);
int gptst_cap_recon_hop(const float* c, const float* s, const float* dyn, const float* e1, float* v, float* recon, int B,
                        int T, int N, int D, int H, int HT, void* stream);
/* recon_hop with hop_e1 folded in (every slab CTA recomputes its sample's E1 from s, 123k MACs): e1 is an OUTPUT here, kept for
 * the backward pass.  Opt-in (GPTST_B200_HOP=fused): measured slower than the split pair on the B200 (80 us vs 6.5 + 22 us).   */
int gptst_cap_recon_hop_fused(const float* c, const float* s, const float* dyn, float* e1, float* v, float* recon, int B, int T,
                              int N, int D, int H, int HT, void* stream);
/* ---- cap backward pieces ------------------------------------------------------
 * dv = c drecon, dcr = v drecon^T                                                                            */
int gptst_cap_dv_dcr(const float* c, const float* v, const float* drecon, float* dv, float* dcr, int B, int T, int N,
                     int D, int H, void* stream);
/* dv -> ds (B,T,H,D) and ddyn (B,HT,T*H)                                                                      */
int gptst_cap_hop_bwd(const float* s, const float* dyn, const float* dv, float* ds, float* ddyn, int B, int T, int D,
                      int H, int HT, void* stream);
/* the same, split for parallelism (per-slab row pass + per-(sample, 16-column) pass); e1 is the tensor hop_e1 wrote in
 * forward; dr_tmp / dpre2_tmp are (B,T,H,D) scratch; ddyn_part is (gptst_cap_hop_bwd_parts(D), B, HT, T*H), the caller sums it. */
int gptst_cap_hop_bwd_parts(int D);
int gptst_cap_hop_bwd2(const float* s, const float* dyn, const float* e1, const float* dv, float* dr_tmp, float* dpre2_tmp,
                       float* ds, float* ddyn_part, int B, int T, int D, int H, int HT, void* stream);
/* dv/dcr fused with the row pass of the hop backward (D = 64, N <= 256): writes dcr and, instead of dv, dr = squash'(r, dv) and
 * dpre2 = dr * phi'(pre2); gptst_cap_hop_bwd_cols is the remaining column pass (-> ds, ddyn_part).                       */
int gptst_cap_dv_dcr_hoprows(const float* c, const float* v, const float* drecon, const float* s, const float* dyn,
                             const float* e1, float* dcr, float* dr, float* dpre2, int B, int T, int N, int D, int H, int HT,
                             void* stream);
int gptst_cap_hop_bwd_cols(const float* s, const float* dyn, const float* e1, const float* dr, const float* dpre2, float* ds,
                           float* ddyn_part, int B, int T, int D, int H, int HT, void* stream);
/* dx_io holds dy = dOut*act'(out) on entry and dy + dZ Wp on exit; ddadj (B,T,H,N);
 * dWp_part (parts,D,D) [out][in], dbp_part (parts,D) with parts = gptst_cap_route_bwd_parts(...)              */
int gptst_cap_route_bwd_parts(int B, int T, int N, int D, int H);
int gptst_cap_route_bwd(const float* x, const float* Wp, const float* bp, const float* c, const float* ds,
                        const float* dcr, float* dx_io, float* ddadj, float* dWp_part, float* dbp_part, int B, int T,
                        int N, int D, int H, int prec, void* stream);

/* ---- cap backward through the routing block, second generation (D = 64, N <= 256; gptst_cap_route2_supported tells) ---
 * route_bwd_dz: recompute Z = x Wp^T + bp, P = squash(Z); dc = dcr + ds P^T; ddadj = c*(dc - sum_h c dc); dP = c^T ds;
 *               dZ = squash'(Z, dP) -> dZ (B,T,N,D).  The two contractions with the shared weight follow as
 * linear_bwd_acc: dX_io += dY W ; dW_part[s] = partial dY^T X ([out][in]) ; db_part[s] = partial sum dY, on (rows, D)
 *               row-major operands; splits = gptst_linear_bwd_acc_splits(rows, D) partials, summed by the caller.      */
int gptst_cap_route2_supported(int N, int D, int H);
int gptst_cap_route_bwd_dz(const float* x, const float* Wp, const float* bp, const float* c, const float* ds,
                           const float* dcr, float* dZ, float* ddadj, int B, int T, int N, int D, int H, int prec,
                           void* stream);
/* Training pair that trades one activation for the recompute (reference GPTST.py:102-103 forward, SURVEY.md appendix A backward):
 * route_fwd_z = gptst_cap_route_fwd that also stores Z = x Wp^T + bp (B,T,N,D); route_bwd_dz_z = route_bwd_dz reading that Z
 * (no x / Wp staging, no Z product).  Same supported geometries as gptst_cap_route2_supported.                              */
int gptst_cap_route_fwd_z(const float* x, const float* Wp, const float* bp, const float* dadj, float* c, float* s, float* z,
                          int B, int T, int N, int D, int H, int R, int prec, void* stream);
int gptst_cap_route_bwd_dz_z(const float* z, const float* c, const float* ds, const float* dcr, float* dZ, float* ddadj,
                             int B, int T, int N, int D, int H, int prec, void* stream);
int gptst_linear_bwd_acc_splits(long rows, int D);
int gptst_linear_bwd_acc(const float* dY, const float* X, const float* W, float* dX_io, float* dW_part, float* db_part,
                         long rows, int D, int prec, int splits, void* stream);

/* ---- on-device mask construction (GPTST.py:312-413) ------------------------------------------
 * mask_labels: label[i] = argmax_h prob[i][h] (uint8, first maximum), counts[h] = class histogram (int32, zeroed here).
 * mask_select: mode 2 (adaptive phase): plan = {shuffled class order[0..H), adaptive_num, random_num} (int64), u1 / u2 the two
 *              torch.rand draws of the reference, m_ada n bytes of scratch; classes are taken in `order` until the adaptive
 *              budget is covered (all but the last masked outright when all_type != 0), the last one is sub-sampled by an exact
 *              top-k of u1 (ties in index order, as the reference's stable sort), then random_num more cells by a top-k of u2.
 *              mode 1 (random phase): plan[0] = number of cells to mask by a top-k of u1.
 *              final_mask (n, i0) int64: 1 = keep, 0 = masked.  One CTA, no host synchronisation.                       */
int gptst_mask_labels(const float* prob, unsigned char* label, int* counts, long n, int H, void* stream);
int gptst_mask_select(const unsigned char* label, const int* counts, const long long* plan, const float* u1, const float* u2,
                      unsigned char* m_ada, long long* final_mask, long n, int H, int i0, int all_type, int mode, void* stream);

/* The same masks through a multi-CTA pipeline (histogram / collect / apply passes per selection; this is what the model
 * launches -- the one-CTA kernel above is its exact slow path and specification).  ws: gptst_mask_ws_ints() int32 scratch;
 * label, m_ada: n bytes scratch each; label_in != NULL supplies the class labels instead of arg-max(prob).                 */
int gptst_mask_ws_ints(void);
int gptst_mask_adaptive(const float* prob, const unsigned char* label_in, const long long* plan, const float* u1,
                        const float* u2, unsigned char* label, unsigned char* m_ada, int* ws, long long* final_mask, long n,
                        int H, int i0, int all_type, void* stream);
int gptst_mask_random(const long long* k_dev, const float* u, int* ws, long long* final_mask, long n, void* stream);

/* ---- small parameter-side ops that sat on the critical tail of the captured step as chains of tiny library kernels -------
 * time_mlp: the time-embedding MLP of GPTST.py:187-219 (h0 = a Wd^T + bd + b Ww^T + bw; z1 = h0 W1^T + b1; z2 = relu(z1) W2^T + b2;
 * out = relu(z2) W3^T + b3) on R rows with F inputs per branch (row stride in_stride) and e <= 16 units; weights [out][in].
 * backward writes (gptst_time_mlp_chunks(R), gptst_time_mlp_grad_floats(e,F)) partials packed as
 * dW3 | db3 | dW2 | db2 | dW1 | db1 | dWd | dbd | dWw | dbw, summed by the caller.
 * affine1_bwd: y = x w + b with one input feature (GPTST.py:22, :298): part[p][0] = partial sum_i dy[i,:] x[i], part[p][1] = partial sum_i dy[i,:]. */
int gptst_time_mlp_chunks(int R);
int gptst_time_mlp_grad_floats(int e, int F);
int gptst_time_mlp_fwd(const float* a, const float* b, const float* Wd, const float* bd, const float* Ww, const float* bw,
                       const float* W1, const float* b1, const float* W2, const float* b2, const float* W3, const float* b3,
                       float* h0, float* z1, float* z2, float* out, int R, int F, int e, long in_stride, void* stream);
int gptst_time_mlp_bwd(const float* a, const float* b, const float* W1, const float* W2, const float* W3, const float* h0,
                       const float* z1, const float* z2, const float* g, float* part, int R, int F, int e, long in_stride,
                       void* stream);
/* forward of a low-rank table: tab (G,C) = te (G,d) . pool (d,C), d <= 16 */
int gptst_table_fwd(const float* te, const float* pool, float* tab, int G, int d, int C, void* stream);
/* backward of a low-rank table Tab = te . pool (te (G,d), pool (d,C), d <= 16; GPTST.py:104, :129, :137-138, :160-161, :24-31):
 * dpool[k][c] = sum_g te[g][k] dTab[g][c], dte[g][k] = sum_c dTab[g][c] pool[k][c]; either output may be NULL.          */
int gptst_table_bwd(const float* te, const float* pool, const float* dtab, float* dpool, float* dte, int G, int d, int C,
                    void* stream);
/* the same, fused on the tensor cores (dTab read once): dpool_part (row_chunks, 16, C), dte_part (col_chunks, G, 16), chunk
 * counts from gptst_table_bwd2_chunks; the caller sums the partials and drops the rows / columns >= d.                     */
int gptst_table_bwd2_chunks(int G, int C, int* row_chunks, int* col_chunks);
int gptst_table_bwd2(const float* te, const float* pool, const float* dtab, float* dpool_part, float* dte_part, int G, int d,
                     int C, void* stream);
/* per-node mix matrix of hyperTem, GPTST.py:156-158: M[n] = A[n]^T A[n] (A (N,Ht,T)) and its backward dA = A (dM + dM^T)   */
int gptst_mn_fwd(const float* A, float* M, int N, int Ht, int T, void* stream);
int gptst_mn_bwd(const float* A, const float* dM, float* dA, int N, int Ht, int T, void* stream);
/* score head of the mask scorer (GPTST.py:33 + softmax of :332/:343): prob (rows,H) = softmax(h (rows,D) W3^T + b3), W3 (H,D)  */
int gptst_score_head_fwd(const float* h, const float* W3, const float* b3, float* prob, long rows, int D, int H, void* stream);
/* sums the (parts[k], numel[k]) partial buffers of up to 8 tensors in one launch, fixed order: outs[k][i] = sum_p ins[k][p][i]
 * (host arrays of device pointers / sizes; the values are copied into the launch arguments)                                */
/* backward of the score head: dh (rows,D) = dz W3 with dz = prob*(dprob - <prob,dprob>); part (gptst_score_head_bwd_parts(rows),
 * H*D + H) = per-CTA partials of dW3 = dz^T h and db3 = sum dz, summed by the caller; dh may be NULL                        */
int gptst_score_head_bwd_parts(long rows);
int gptst_score_head_bwd(const float* h, const float* W3, const float* prob, const float* dprob, float* dh, float* part,
                         long rows, int D, int H, void* stream);
int gptst_sum_partials(const float* const* ins, float* const* outs, const long* numel, const int* parts, int n, void* stream);
int gptst_affine1_fwd(const float* x, const float* w, const float* b, float* y, long n, int D, void* stream);
/* masked input embedding of the encoder, reference GPTST.py:419-421 (torch.where(mask == 0, scaler_zeros, mask * flow) followed by
 * dim_in_flow): xm[i] = mask[i] == 0 ? fill : mask[i] * flow[i * flow_stride]; y[i][:] = xm[i] w[:] + b[:]; mask int64 (n,) */
int gptst_masked_affine1_fwd(const float* flow, long flow_stride, const long long* mask, float fill, const float* w,
                             const float* b, float* xm, float* y, long n, int D, void* stream);
int gptst_affine1_bwd_parts(long n);
int gptst_affine1_bwd(const float* dy, const float* x, float* part, long n, int D, int parts, void* stream);
/* Sign-mask backward of gptst_gproj_fwd for D = 64 (csrc/gproj3.cu): dy = dY * act'(mask) with mask = 8 bytes per row of Y
 * (bit c = Y[c] > 0, rows in Y's memory order).  flags: bit 0 = dX accumulated in place, bit 1 = W / dW are [out][in].
 * Outputs as gptst_gproj_bwd.  (The default path reaches this kernel through gptst_hypertem_dw.)                              */
int gptst_gproj3_bwd(const float* dY, const void* mask, const float* X, const float* W, float* dX, float* dW_part,
                     float* dbias_part, float* dRes, int G, int R, long group_stride, long row_stride, int D, int act, int prec,
                     int splits, int flags, void* stream);
/* decoder output projection dim_flow_out = nn.Linear(D, O), O = input_base_dim <= 4 (GPTST.py:454-458), replacing
 * `self.dim_flow_out(flow_decode)` and its autograd: y (rows,O) = x (rows,D) W^T + b, W (O,D) as nn.Linear stores it, D = 64|128.
 * Backward in one pass: dX (rows,D) = dy W (may be NULL) and part (parts, O*D + O) = per-CTA partials of dW = dy^T x
 * ([p][o*D + d]) and db = sum dy ([p][O*D + o]), parts = gptst_proj_out_bwd_parts(rows), summed by the caller.              */
int gptst_proj_out_fwd(const float* x, const float* W, const float* b, float* y, long rows, int D, int O, void* stream);
int gptst_proj_out_bwd_parts(long rows);
int gptst_proj_out_bwd(const float* dy, const float* x, const float* W, float* dX, float* part, long rows, int D, int O,
                       int parts, void* stream);

/* ---- fused pre-training loss + analytic gradients ------------------------------------
 * mode 0: probe loss mean|(o - x)*m| ; mode 1: masked MAE of Run.py:91-101 / lib/metrics.py:11-18 (inverse z-score with
 * mean/std, keep true*m > thr) ; plus kl_w * KLDivLoss(sum)(log prob, hs) (BasicTrainer.py:84-86) when kl_w != 0.
 * o (cells, ibd) fp32, src (cells, src_stride) fp32 (first ibd channels are the flow), inv_mask (cells, ibd) int64,
 * prob / hs (cells, H).  Outputs: d_o = dloss/do, d_prob = dloss/dprob, out = {loss, mae, kl_sum};
 * part = scratch of 3*gptst_loss_parts() floats.  Two stream-ordered launches, deterministic.                  */
int gptst_loss_parts(void);
int gptst_pretrain_loss(const float* o, const float* src, const long long* inv_mask, const float* prob, const float* hs,
                        float* d_o, float* d_prob, float* part, float* out, long n_cells, int ibd, int src_stride, int H,
                        int mode, float mean, float std_, float thr, float kl_w, void* stream);

/* ---- fused global-norm clipping + Adam over a tensor list (BasicTrainer.py:94-97, Run.py:134) --------------
 * table: device array of 6 x int64 per tensor {param ptr, grad ptr, exp_avg ptr, exp_avg_sq ptr, numel, first_step};
 * block_map: device int2 per block {tensor index, chunk index} with gptst_opt_chunk() elements per chunk;
 * partial: nblocks floats scratch; step: device int32 global step (incremented here); hyper: device float[5]
 * {lr, beta1, beta2, eps, max_norm (<= 0 disables clipping)}; norm_out: optional device float (pre-clip grad norm).   */
int gptst_opt_chunk(void);
int gptst_adam_clip(const void* table, const void* block_map, int nblocks, float* partial, int* step, const float* hyper,
                    float* norm_out, void* stream);

/* ---- cap forward, round 2 (reference GPTST.py:125-141, D = 64, three-term split) ------------------------------------------
 * gptst_cap_hop_ev: the whole inter-cluster hop (GPTST.py:125-134) in ONE launch, one CTA per sample:
 *   e1 (B,HT,D) = LReLU(dyn_b (s_b + tau)),  v (B,T,H,D) = squash(LReLU(dyn_b^T e1) + s);  s (B,T,H,D), dyn (B,HT,T*H).
 *   Replaces gptst_cap_hop_e1 + the first half of gptst_cap_recon_hop.
 * gptst_cap_recon_proj = GPTST.py:135-141 in ONE launch: out = LReLU((c^T v) W_n + bias_n + x), W_n given as the fragment table
 *   gptst_hypertem_pack_w(W_n, wfrag, NULL, N) writes (gptst_hypertem_wfrag_bytes(N) bytes); recon (B,T,N,D) = c^T v is stored
 *   when non-NULL (the backward reads it).  Replaces the second half of gptst_cap_recon_hop + the node-grouped gptst_gproj_fwd:
 *   `recon` no longer makes a round trip through L2 / HBM between the two.
 * Both return -2 outside the geometry they cover (the caller then uses the older chain).                                       */
int gptst_cap_hop_ev(const float* s, const float* dyn, float* e1, float* v, int B, int T, int D, int H, int HT, void* stream);
int gptst_cap_recon_proj(const float* c, const float* v, const float* x, const void* wfrag, const float* bias, float* out,
                         float* recon, int B, int T, int N, int D, int H, void* stream);

/* ---- eval-path glue next to the encoder ----------------------------------------------
 * Fusion gate of Enhance_model (reference model/Model.py:12-17, called from Model.py:106-109):
 *     z = sigmoid(HS_fc(flow) + HT_fc(time)),  h = z * flow + (1 - z) * time          (then output_fc(h))
 * gptst_gate_fwd replaces lines 13-16 with ONE launch: the second product HT_fc(time) with the sigmoid and the blend as its
 * epilogue (D = 64, three-term split); xs = HS_fc(flow) comes from a plain gptst_gproj_fwd, W is HT_fc.weight^T ([in][out]),
 * z (rows, D) is kept for the backward when non-NULL.  gptst_gate_blend is the same gate as an elementwise kernel on the two
 * pre-activations (any width; n elements, n % 4 == 0); gptst_gate_bwd gives dpre = dh (flow - time) z (1 - z) (the gradient of both
 * pre-activations), dx = dh z (direct path into flow) and dy = dh (1 - z) (direct path into time).                               */
int gptst_gate_fwd(const float* flow, const float* time, const float* xs, const float* W, const float* bias, float* h, float* z,
                   long rows, int D, int prec, void* stream);
int gptst_gate_blend(const float* xs, const float* xt, const float* x, const float* y, float* h, float* z, long n, void* stream);
int gptst_gate_bwd(const float* dh, const float* z, const float* x, const float* y, float* dpre, float* dx, float* dy, long n,
                   void* stream);
/* STGCN's gated (GLU) temporal convolution, reference model/STGCN/stgcn.py:25-53 (TemporalConvLayer, act = "GLU") with the
 * Align of stgcn.py:10-23 folded in:  out = (conv(x)[:, :Cout] + align(x)) * sigmoid(conv(x)[:, Cout:]),
 * conv = Conv2d(Cin, 2 Cout, (kt, 1), padding (kt-1)/2).  x (B, Cin, T, N) and out (B, Cout, T, N) in the reference's own layout
 * (N fastest), W (2 Cout, Cin, kt) = conv.weight without its trailing 1, bias (2 Cout); aw (Cout, Cin) / ab (Cout) = the 1x1
 * Align conv, only when Cin > Cout (NULL otherwise; Cin < Cout zero-pads the channels, Cin == Cout is the identity).
 * kt odd (an even kt shortens the sequence and the reference's own residual add fails), T <= 12.  P = conv[:, :Cout] + align(x)
 * and S = sigmoid(conv[:, Cout:]) are stored for the backward when both are non-NULL.
 * Backward pieces: gptst_glu_gate_bwd: dconv (B, 2 Cout, T, N) = [dout S ; dout P S (1 - S)];
 * gptst_tconv_fwd: plain "same"-padded temporal convolution (used for dx = conv of dconv with the flipped, transposed weights,
 * the Align term folded into the centre tap by the caller); gptst_tconv_dw: per-split partials dW_part (splits, C2, Cin, kt),
 * db_part (splits, C2) of dW[o,i,k] = sum dconv[b,o,t,n] x[b,i,t+k-pad,n], db[o] = sum dconv, splits = gptst_tconv_dw_splits().   */
int gptst_glu_tconv_fwd(const float* x, const float* W, const float* bias, const float* aw, const float* ab, float* out, float* P,
                        float* S, int B, int Cin, int Cout, int T, int N, int kt, void* stream);
int gptst_tconv_fwd(const float* x, const float* W, const float* bias, float* out, int B, int Cin, int Cout, int T, int N, int kt,
                    void* stream);
int gptst_glu_gate_bwd(const float* dout, const float* P, const float* S, float* dconv, int B, int Cout, int T, int N, void* stream);
int gptst_tconv_dw_splits(int B, int C2, int Cin, int N);
int gptst_tconv_dw(const float* dconv, const float* x, float* dW_part, float* db_part, int B, int C2, int Cin, int T, int N, int kt,
                   int splits, void* stream);

/* library identification: "gptst_b200 <version> sm_100a" */
const char* gptst_version(void);

#ifdef __cplusplus
}
#endif
#endif /* GPTST_B200_H */
