// bf16 tensor-core versions of the small dense layers around the SIM token kernels
// (q / folded-query prep, value + output projections, FFN; useA.py:388-401) -- included by sim.cu.
// Activations keep an fp32 master (LayerNorm, residuals, softmax statistics run in fp32) and a bf16
// shadow that feeds the next tcgen05 GEMM; weights are cast to bf16 once per call.

static TcGemmDesc lin_nt(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* W, int64_t ldw, float* C, int64_t ldc,
                         const float* bias, int M, int N, int K) {
  TcGemmDesc t = tc_desc();
  t.A = tc_k2d(A, M, K, lda);
  t.B = tc_k2d(W, N, K, ldw);
  t.M = M; t.N = N; t.K = K;
  t.C[0] = C; t.ldc = ldc; t.bias[0] = bias;
  return t;
}
// C[M,N] = A[M,K] Wm[K,N]
static TcGemmDesc lin_nn(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* Wm, int64_t ldw, float* C, int64_t ldc, int M,
                         int N, int K) {
  TcGemmDesc t = tc_desc();
  t.A = tc_k2d(A, M, K, lda);
  t.B = tc_mn2d(Wm, K, N, ldw);
  t.M = M; t.N = N; t.K = K;
  t.C[0] = C; t.ldc = ldc;
  return t;
}
// C[M,N] = A[K,M]^T Bm[K,N]
static TcGemmDesc lin_tn(const __nv_bfloat16* A, int64_t lda, const __nv_bfloat16* Bm, int64_t ldb, float* C, int64_t ldc, int M,
                         int N, int K) {
  TcGemmDesc t = tc_desc();
  t.A = tc_mn2d(A, K, M, lda);
  t.B = tc_mn2d(Bm, K, N, ldb);
  t.M = M; t.N = N; t.K = K;
  t.C[0] = C; t.ldc = ldc;
  return t;
}
// per-head batch: entry h uses pointers base + h * stride (elements)
static void per_head(TcOperand& o, const __nv_bfloat16* base, int64_t stride) {
  for (int h = 0; h < kHeads; ++h) o.ptr[h] = base + h * stride;
}

// The dense layers of the chain have M = 3B rows: at B = 128 that is 3 M-tiles x 6..12 N-tiles = 18..36 work units, each
// pulling its whole K range (393 KB at K = 768) through ONE SM's ~70 GB/s L2 port -- 4-6 us per GEMM while 110+ SMs idle.
// Split-K spreads the same bytes over ~148 units: fp32 partial tiles are added into a pre-zeroed C (red.global.add.v4,
// the unit with the first k-blocks also adds the bias).  Only for GEMMs whose consumer is an element-wise kernel reading the
// fp32 C (LayerNorm, GELU + shadow) or that accumulate onto existing values.
// MEASURED (B = 128, d = 768, one B200, profiles/r2_splitk_experiment.json): the phases shrink where the GEMM dominated
// (sim_post 56.4 -> 52.3 us, sim_attn_prep_bwd 28.1 -> 21.3) but sim_post_bwd grows (88.8 -> 93.6) and the fused step does not
// improve (0.6014 vs 0.5878 ms): next to AlignM's GPU-filling kernels the extra CTAs and 7 MB of atomics per GEMM cost what
// the shorter K loops save.  Off by default; SIG_SIM_SPLITK=1 turns it on.
static bool sim_splitk_enabled() {
  static const bool on = [] {
    const char* e = getenv("SIG_SIM_SPLITK");
    return e && e[0] == '1';
  }();
  return on;
}
static void split_k(TcGemmDesc& t) {
  if (!sim_splitk_enabled() || t.out_bf16 || t.act || t.C2[0] || t.pre[0] || t.batch != 1) return;
  const int tiles = (int)(ceil_div(t.M, 128) * ceil_div(t.N, t.bn));
  const int kblocks = (int)ceil_div(t.K, 64);
  int ks = (tc_num_sms() + tiles - 1) / tiles;
  if (ks > kblocks / 2) ks = kblocks / 2;       // at least two k-blocks per unit
  if (ks < 2) return;
  t.ksplit = ks;
  t.accumulate = 0;                              // atomics add onto what is there: that IS the accumulation
}

// c[(b,q),h] = scale * q_h . b_k^h   grid R, 256 threads (one warp per head)
static __global__ void __launch_bounds__(256) catt_kernel(const float* __restrict__ qatt, const float* __restrict__ bk, int d, float scale,
                                                          float* __restrict__ catt) {
  pdl_enter();
  const int64_t row = blockIdx.x;
  const int h = threadIdx.x >> 5, lane = threadIdx.x & 31, hd = d / kHeads;
  float a = 0.f;
  for (int c = lane; c < hd; c += 32) a = fmaf(qatt[row * d + h * hd + c], bk[h * hd + c], a);
  a = warp_sum(a);
  if (lane == 0) catt[row * 8 + h] = a * scale;
}

// shadow = bf16(gelu(a)), 4 elements per thread (n is a multiple of 4: rows of 2d)
static __global__ void gelu_shadow_kernel(const float* __restrict__ a, __nv_bfloat16* __restrict__ shadow, int64_t n) {
  pdl_enter();
  const int64_t i = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(a + i);
    const float t[4] = {gelu_f(v.x), gelu_f(v.y), gelu_f(v.z), gelu_f(v.w)};
    store4(shadow + i, t);
  }
}

// da = dh * gelu'(a) in place (fp32) + bf16 shadow
static __global__ void gelu_bwd_shadow_kernel(float* dh, const float* __restrict__ a, __nv_bfloat16* __restrict__ shadow, int64_t n) {
  pdl_enter();
  const int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x;
  if (i < n) {
    const float v = dh[i] * gelu_grad_f(a[i]);
    dh[i] = v;
    shadow[i] = __float2bfloat16_rn(v);
  }
}

static int cast_sim_weights(const SimCtx& c, const sig_sim_params* p, int d, cudaStream_t s) {
  const int64_t dd = (int64_t)d * d;
  const CastJob jobs[4] = {{p->in_proj_w, c.Wb, 3 * dd}, {p->out_proj_w, c.Wb + 3 * dd, dd}, {p->ffn0_w, c.Wb + 4 * dd, 2 * dd},
                           {p->ffn2_w, c.Wb + 6 * dd, 2 * dd}};
  return cast_f32_to_bf16_multi(jobs, 4, s);
}

static int attn_prep_tc(const SimCtx& c, const sig_sim_params* p, int B, int d, cudaStream_t s) {
  const int R = 3 * B, hd = d / kHeads;
  const float scale = 1.0f / sqrtf((float)hd);
  const size_t dd = (size_t)d * d;
  const __nv_bfloat16* wqb = c.Wb;
  const __nv_bfloat16* wkb = c.Wb + dd;
  if (sim_splitk_enabled())   // [attn | a1 | f]: split-K targets of attn_post_tc (one memset, far ahead of their GEMMs)
    cudaMemsetAsync(c.attn, 0, (size_t)(reinterpret_cast<char*>(c.f) - reinterpret_cast<char*>(c.attn)) + (size_t)R * d * sizeof(float), s);
  SIG_TRY(cast_sim_weights(c, p, d, s));
  {  // q = W_q cls + b_q  (fp32 + bf16 shadow)
    TcGemmDesc t = lin_nt(c.clsb, d, wqb, d, c.qatt, d, p->in_proj_b, R, d, d);
    t.C2[0] = c.qattb; t.ldc2 = d;
    SIG_TRY(tc_gemm(t, s));
  }
  {  // qt[(b,q),h,:] = scale * q_h W_k^h
    TcGemmDesc t = lin_nn(c.qattb, d, wkb, d, c.qtatt, 8 * (int64_t)d, R, d, hd);
    t.batch = kHeads; t.alpha = scale;
    per_head(t.A, c.qattb, hd);
    per_head(t.B, wkb, (int64_t)hd * d);
    for (int h = 0; h < kHeads; ++h) t.C[h] = c.qtatt + (size_t)h * d;
    SIG_TRY(tc_gemm(t, s));
  }
  SIG_LAUNCH((catt_kernel), R, 256, 0, s, c.qatt, p->in_proj_b + d, d, scale, c.catt);
  SIG_CHECK_LAUNCH();
  return 0;
}

template <typename OutT>
static int attn_post_tc(const SimCtx& c, const sig_sim_params* p, int B, int d, OutT* out, cudaStream_t s) {
  const int R = 3 * B, hd = d / kHeads;
  const size_t dd = (size_t)d * d;
  const __nv_bfloat16* wvb = c.Wb + 2 * dd;
  const __nv_bfloat16* wob = c.Wb + 3 * dd;
  const __nv_bfloat16* w1b = c.Wb + 4 * dd;
  const __nv_bfloat16* w2b = c.Wb + 6 * dd;
  // (xbarb, the bf16 shadow of xbar, was written by the pooling kernel's epilogue)
  {  // o_h = W_v^h xbar_h + b_v^h
    TcGemmDesc t = lin_nt(c.xbarb, 8 * (int64_t)d, wvb, d, c.o, d, nullptr, R, hd, d);
    t.batch = kHeads;
    per_head(t.A, c.xbarb, d);
    per_head(t.B, wvb, (int64_t)hd * d);
    t.ldc2 = d;
    for (int h = 0; h < kHeads; ++h) {
      t.C[h] = c.o + (size_t)h * hd;
      t.bias[h] = p->in_proj_b + 2 * d + h * hd;
      t.C2[h] = c.ob + (size_t)h * hd;
    }
    SIG_TRY(tc_gemm(t, s));
  }
  {
    TcGemmDesc t = lin_nt(c.ob, d, wob, d, c.attn, d, p->out_proj_b, R, d, d);
    split_k(t);
    SIG_TRY(tc_gemm(t, s));
  }
  SIG_LAUNCH((layernorm_fwd_kernel<float>), R, 256, 0, s, c.attn, c.clsf, p->ln1_w, p->ln1_b, d, c.r1, c.mu1, c.rstd1, c.y1, c.y1b);
  SIG_CHECK_LAUNCH();
  {
    // a1 = y1 W1^T + b1 (fp32, kept for the backward), then h1 = gelu(a1) as the bf16 operand of the next GEMM.
    // The activation is a separate full-grid elementwise launch: inside the GEMM it would run on the four
    // epilogue warps of 36 CTAs (erf on one warp per scheduler), which measured 3x the GEMM itself.
    TcGemmDesc t = lin_nt(c.y1b, d, w1b, d, c.a1, 2 * (int64_t)d, p->ffn0_b, R, 2 * d, d);
    split_k(t);
    SIG_TRY(tc_gemm(t, s));
    const int64_t n = (int64_t)R * 2 * d;
    SIG_LAUNCH((gelu_shadow_kernel), (unsigned)ceil_div(n, 4 * 256), 256, 0, s, c.a1, c.h1b, n);
    SIG_CHECK_LAUNCH();
  }
  {
    TcGemmDesc t = lin_nt(c.h1b, 2 * (int64_t)d, w2b, 2 * (int64_t)d, c.f, d, p->ffn2_b, R, d, 2 * d);
    split_k(t);
    SIG_TRY(tc_gemm(t, s));
  }
  SIG_LAUNCH((layernorm_fwd_kernel<OutT>), R, 256, 0, s, c.f, c.y1, p->ln2_w, p->ln2_b, d, c.r2, c.mu2, c.rstd2, out, (__nv_bfloat16*)nullptr);
  SIG_CHECK_LAUNCH();
  return 0;
}

// Weight-gradient GEMMs only read bf16 shadows that are never rewritten inside the call, so they run on
// the side stream `fk` next to the activation-gradient chain (forked after their operands exist).
static int side_gemm(const Fork& fk, cudaStream_t s, const TcGemmDesc& t) {
  if (!fk.ok()) return tc_gemm(t, s);
  fk.fork(s);
  return tc_gemm(t, fk.side);
}

// Column sums of buffers that are not written again inside the call: next to the weight-gradient GEMMs.
static int side_colsum(const Fork& fk, cudaStream_t s, const float* X, int64_t ldx, int M, int N, float* out) {
  if (!fk.ok()) return launch_colsum(X, ldx, M, N, out, 1.f, s);
  fk.fork(s);
  return launch_colsum(X, ldx, M, N, out, 1.f, fk.side);
}

template <typename InT>
static int attn_post_bwd_tc(const SimCtx& c, const sig_sim_params* p, int B, int d, const InT* dout, const sig_sim_param_grads* g,
                            const Fork& fk, cudaStream_t s) {
  const int R = 3 * B, hd = d / kHeads;
  const size_t dd = (size_t)d * d;
  const __nv_bfloat16* wvb = c.Wb + 2 * dd;
  const __nv_bfloat16* wob = c.Wb + 3 * dd;
  const __nv_bfloat16* w1b = c.Wb + 4 * dd;
  const __nv_bfloat16* w2b = c.Wb + 6 * dd;
  float* dwv = g->in_proj_w + 2 * dd;
  // the bf16 weight shadows written by the forward call are reused: autograd's version check on the
  // saved parameters rules out an in-place update between forward and backward
  if (sim_splitk_enabled())   // split-K target of the dh1 GEMM below (ahead of the kernel chain, not inside it)
    cudaMemsetAsync(c.dh1, 0, (size_t)R * 2 * d * sizeof(float), s);
  SIG_LAUNCH((layernorm_bwd_kernel<InT>), R, 256, 0, s, dout, c.r2, p->ln2_w, c.mu2, c.rstd2, nullptr, d, c.dr2, c.dyx, c.dyf, c.dr2b);
  SIG_CHECK_LAUNCH();
  // (the three column sums read buffers that are rewritten further down this stream -- dyx/dyf by the second
  //  LayerNorm backward, dr2 by the accumulating GEMM -- so they stay on it, as ONE launch)
  SIG_TRY(launch_colsum3(c.dyx, g->ln2_w, c.dyf, g->ln2_b, c.dr2, g->ffn2_b, d, R, d, s));
  SIG_TRY(side_gemm(fk, s, lin_tn(c.dr2b, d, c.h1b, 2 * (int64_t)d, g->ffn2_w, 2 * (int64_t)d, d, 2 * d, R)));   // dW2 = dr2^T h1
  {  // dh1 = dr2 W2
    TcGemmDesc t = lin_nn(c.dr2b, d, w2b, 2 * (int64_t)d, c.dh1, 2 * (int64_t)d, R, 2 * d, d);
    split_k(t);      // (c.dh1 was zeroed at the top of this function)
    SIG_TRY(tc_gemm(t, s));
  }
  {
    const int64_t n = (int64_t)R * 2 * d;
    SIG_LAUNCH((gelu_bwd_shadow_kernel), (unsigned)ceil_div(n, 256), 256, 0, s, c.dh1, c.a1, c.da1b, n);
    SIG_CHECK_LAUNCH();
  }
  SIG_TRY(side_colsum(fk, s, c.dh1, 2 * (int64_t)d, R, 2 * d, g->ffn0_b));   // dh1 is not written again
  SIG_TRY(side_gemm(fk, s, lin_tn(c.da1b, 2 * (int64_t)d, c.y1b, d, g->ffn0_w, d, 2 * d, d, R)));                 // dW1 = da1^T y1
  {  // dy1 = dr2 + da1 W1
    TcGemmDesc t = lin_nn(c.da1b, 2 * (int64_t)d, w1b, d, c.dr2, d, R, d, 2 * d);
    t.accumulate = 1;
    split_k(t);
    SIG_TRY(tc_gemm(t, s));
  }
  SIG_LAUNCH((layernorm_bwd_kernel<float>), R, 256, 0, s, c.dr2, c.r1, p->ln1_w, c.mu1, c.rstd1, nullptr, d, c.dr1, c.dyx, c.dyf, c.dr1b);
  SIG_CHECK_LAUNCH();
  SIG_TRY(launch_colsum3(c.dyx, g->ln1_w, c.dyf, g->ln1_b, c.dr1, g->out_proj_b, d, R, d, s));   // (dr1 is accumulated into later)
  SIG_TRY(side_gemm(fk, s, lin_tn(c.dr1b, d, c.ob, d, g->out_proj_w, d, d, d, R)));                                // dWo = dr1^T o
  {  // do = dr1 Wo
    TcGemmDesc t = lin_nn(c.dr1b, d, wob, d, c.dob, d, R, d, d);
    t.C2[0] = c.dobb; t.ldc2 = d;
    SIG_TRY(tc_gemm(t, s));
  }
  SIG_TRY(side_colsum(fk, s, c.dob, d, R, d, g->in_proj_b + 2 * d));
  {  // dW_v^h = do_h^T xbar_h
    TcGemmDesc t = lin_tn(c.dobb, d, c.xbarb, 8 * (int64_t)d, dwv, d, hd, d, R);
    t.batch = kHeads;
    per_head(t.A, c.dobb, hd);
    per_head(t.B, c.xbarb, d);
    for (int h = 0; h < kHeads; ++h) t.C[h] = dwv + (size_t)h * hd * d;
    SIG_TRY(side_gemm(fk, s, t));
  }
  // every parameter gradient except W_q, W_k and in_proj_b is now enqueued: on the side stream, which has
  // waited for this stream's column sums (sig_sim_param_grads.early_event)
  if (g->early_event) cudaEventRecord((cudaEvent_t)g->early_event, fk.ok() ? fk.side : s);
  {  // dxbar_h = do_h W_v^h
    TcGemmDesc t = lin_nn(c.dobb, d, wvb, d, c.dxbar, 8 * (int64_t)d, R, d, hd);
    t.batch = kHeads;
    per_head(t.A, c.dobb, hd);
    per_head(t.B, wvb, (int64_t)hd * d);
    for (int h = 0; h < kHeads; ++h) t.C[h] = c.dxbar + (size_t)h * d;
    SIG_TRY(tc_gemm(t, s));
  }
  return 0;
}

static int attn_prep_bwd_tc(const SimCtx& c, const sig_sim_params* p, int B, int d, const sig_sim_param_grads* g, const Fork& fk,
                            cudaStream_t s) {
  const int R = 3 * B, hd = d / kHeads;
  const float scale = 1.0f / sqrtf((float)hd);
  const size_t dd = (size_t)d * d;
  const __nv_bfloat16* wqb = c.Wb;
  const __nv_bfloat16* wkb = c.Wb + dd;
  float* dwq = g->in_proj_w;
  float* dwk = g->in_proj_w + dd;
  // (dqtb, the bf16 shadow of dqt, was written by the token kernel's epilogue)
  {  // dq_h = scale * dqt_h W_k^hT
    TcGemmDesc t = lin_nt(c.dqtb, 8 * (int64_t)d, wkb, d, c.dqatt, d, nullptr, R, hd, d);
    t.batch = kHeads; t.alpha = scale; t.ldc2 = d;
    per_head(t.A, c.dqtb, d);
    per_head(t.B, wkb, (int64_t)hd * d);
    for (int h = 0; h < kHeads; ++h) {
      t.C[h] = c.dqatt + (size_t)h * hd;
      t.C2[h] = c.dqattb + (size_t)h * hd;
    }
    SIG_TRY(tc_gemm(t, s));
  }
  {  // dW_k^h = scale * q_h^T dqt_h
    TcGemmDesc t = lin_tn(c.qattb, d, c.dqtb, 8 * (int64_t)d, dwk, d, hd, d, R);
    t.batch = kHeads; t.alpha = scale;
    per_head(t.A, c.qattb, hd);
    per_head(t.B, c.dqtb, d);
    for (int h = 0; h < kHeads; ++h) t.C[h] = dwk + (size_t)h * hd * d;
    SIG_TRY(side_gemm(fk, s, t));
  }
  if (fk.ok()) {
    fk.fork(s);
    cudaMemsetAsync(g->in_proj_b + d, 0, d * sizeof(float), fk.side);  // key bias: softmax shift invariance => exactly 0
    SIG_TRY(launch_colsum(c.dqatt, d, R, d, g->in_proj_b, 1.f, fk.side));
  } else {
    cudaMemsetAsync(g->in_proj_b + d, 0, d * sizeof(float), s);
    SIG_TRY(launch_colsum(c.dqatt, d, R, d, g->in_proj_b, 1.f, s));
  }
  SIG_TRY(side_gemm(fk, s, lin_tn(c.dqattb, d, c.clsb, d, dwq, d, d, d, R)));                                       // dWq = dq^T cls
  // every parameter gradient of the call is now enqueued (the late ones on the side stream): sig_sim_param_grads.late_event
  if (g->late_event) cudaEventRecord((cudaEvent_t)g->late_event, fk.ok() ? fk.side : s);
  {  // dcls = dr1 (residual) + dq W_q
    TcGemmDesc t = lin_nn(c.dqattb, d, wqb, d, c.dr1, d, R, d, d);
    t.accumulate = 1;
    split_k(t);
    SIG_TRY(tc_gemm(t, s));
  }
  return 0;
}
