// model.cu - the C ABI (include/kcvae.h) and the host-side orchestration of one
// KurtosisCVAE replica on one B200: topology (src/abstract_cvae.py:22-92), forward
// (:115-149), loss (src/kurtosis_*_cvae.py), hand-derived backward (tape.gradient, :160),
// gradient / moment all-reduce over NCCL, fused Adam (train.py:99-101) and the anomaly
// score (do_anomaly_detection.py:57-117).
#include <cmath>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>

#include "../../include/kcvae.h"
#include "kernels.h"

#ifndef KCVAE_EMU
#include <dlfcn.h>
#include <nccl.h>
#include "tc_conv.h"
#include "tc_gen.h"
#endif

using namespace kc;

// ---- per-launch event profiler (see common.cuh) --------------------------------------------
namespace kc {
const char* g_tag = "";
bool g_prof_on = false;
#ifndef KCVAE_EMU
namespace {
struct ProfRec { std::string key; cudaEvent_t e0, e1; };
std::vector<ProfRec> g_prof_recs;
}
void prof_begin(const char* kernel, cudaStream_t st) {
  ProfRec r;
  r.key = std::string(g_tag) + "/" + kernel;
  cudaEventCreate(&r.e0); cudaEventCreate(&r.e1);
  cudaEventRecord(r.e0, st);
  g_prof_recs.push_back(r);
}
void prof_end(cudaStream_t st) { if (!g_prof_recs.empty()) cudaEventRecord(g_prof_recs.back().e1, st); }
#else
void prof_begin(const char*, cudaStream_t) {}
void prof_end(cudaStream_t) {}
#endif
}  // namespace kc

namespace {

std::string g_create_error;

struct Var {
  int rank;
  int64_t dims[4];
  int64_t off, n;
};

#ifndef KCVAE_EMU
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    // bare soname first: resolves to the copy torch already mapped (same NCCL for both)
    lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) { err = std::string("cannot dlopen libnccl.so.2: ") + dlerror(); return false; }
#define KC_SYM(name) name = reinterpret_cast<decltype(name)>(dlsym(lib, "nccl" #name)); if (!name) { err = "missing nccl" #name; return false; }
    KC_SYM(GetUniqueId) KC_SYM(CommInitRank) KC_SYM(AllReduce) KC_SYM(Broadcast) KC_SYM(CommDestroy) KC_SYM(GetErrorString)
#undef KC_SYM
    return true;
  }
};
NcclApi g_nccl;
#else
typedef void (*emu_allreduce_fn)(void* buf, int64_t count, int dtype, int op);
emu_allreduce_fn g_emu_allreduce = nullptr;
#endif

}  // namespace

struct kcvae_model {
  kcvae_config cfg;
  int device = 0;
  std::string err;
  // topology
  int L = 0, H = 0, W = 0, C = 0, latent = 0, enc_dense = 0, flat = 0, dec_units = 0;
  int64_t P = 0;
  std::vector<int> eh, ew, ec;  // encoder activation sizes, index 0 = input image
  std::vector<int> dh, dw, dc;  // decoder activation sizes, index 0 = Dense output grid
  std::vector<Var> vars;
  int64_t nparams = 0;
  // parameters / optimizer (flat fp32, Keras variable order)
  float *w = nullptr, *g = nullptr, *m = nullptr, *v = nullptr;
  int64_t adam_t = 0;
  float lr = 1e-3f, beta = 0.f;
  LossWeights lw;
  uint64_t seed = 0x5eed5eedULL, seed_base = 0x5eed5eedULL, rng_counter = 0;
  // workspace
  int cap_fwd = 0, cap_bwd = 0, last_B = 0;
  std::vector<float*> act_e, act_d, g_act_e, g_act_d;
  float* x_stage[2] = {nullptr, nullptr};   // double-buffered H2D staging for the *_host entry points
  // bf16 weight images of the tensor-core kernels are rebuilt only after the weights changed
  uint64_t w_version = 1;                   // bumped by every writer of `w`
  uint64_t img_version[5] = {0, 0, 0, 0, 0};   // convT fwd, tail / out conv, out dgrad, convT dgrad, 32 -> few convT fwd
  bool w_external = false;                  // the raw device pointer was handed out: never trust the cache
  kc::ResizePlan* resize_plan = nullptr;    // uint8 front end: span tables of the last (in_h, in_w) seen
  uint8_t* u8_stage[2] = {nullptr, nullptr};   // double-buffered H2D staging of uint8 host frames
  size_t u8_stage_bytes[2] = {0, 0};
  const void* u8_pend_src[2] = {nullptr, nullptr};   // host pointer whose prefetch sits in u8_stage[i]
  size_t u8_pend_bytes[2] = {0, 0};
  cudaEvent_t u8_copy_done[2] = {nullptr, nullptr}, u8_free[2] = {nullptr, nullptr};
  int u8_last = 1;                             // slot the most recent step read
  const void* pend_src[2] = {nullptr, nullptr};   // host pointer whose prefetch sits in x_stage[i] (not yet consumed)
  int pend_batch[2] = {0, 0};
#ifndef KCVAE_EMU
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t copy_done[2] = {nullptr, nullptr}, stage_free[2] = {nullptr, nullptr};
#endif
  float *x_in = nullptr, *d1 = nullptr, *head = nullptr, *z = nullptr, *mean = nullptr, *logvar = nullptr;
  float *eps_buf = nullptr, *xhat = nullptr, *x_noisy = nullptr;
  float *dlogit = nullptr, *g_z = nullptr, *dhead = nullptr, *g_d1 = nullptr;
  float *partial = nullptr, *err_buf = nullptr, *score_buf = nullptr;
  size_t partial_floats = 0;
  double *dpartial = nullptr, *sums = nullptr, *std_acc = nullptr, *pos_sums = nullptr;
  float *minmax = nullptr, *metrics_dev = nullptr;
  // tensor-core path (precision == BF16_TC): bf16 copy of the last decoder activation, UMMA
  // weight image of the output layer, device-side error flag of the bounded barrier waits
  uint32_t* relu_bits = nullptr;  // [B,H,W] ReLU mask of the last activation, one bit per channel (written by the fused tail)
  bool relu_bits_valid = false;
  bool fuse_train_tail = false;  // training forward: fused tail that also stores the activation for the backward
  bool tc_failed = false;        // a tensor-core launcher could not run (tensor map encode): the step is invalid
  bool train_image_noise = false;   // train_step adds N(0, beta^2) to the encoder input (kcvae_set_train_image_noise)
  bool use_tc_out = false, use_tc_dgrad = false, use_tc_convT = false, use_tc_convT_bwd = false;
  void* g_s2d = nullptr;         // bf16 space-to-depth d loss / d a_last, chunk-planar [B][4][4][H/2][W/2][8]
  void* wimg_convT_dgrad = nullptr;
  void* a_prev8 = nullptr;       // bf16 input of the last Conv2DTranspose s2, NHWC padded to 8 channels
  void* a_pp_planar = nullptr;   // chunk-planar bf16 copy of the activation before it (input of the 32 -> few tensor-core layer)
  void* wimg_convT_few = nullptr;
  bool use_tc_convT_few = false;   // inference entry points (scoring, forward, decode)
  bool use_tc_convT_few_train = false;
  void* wimg_convT = nullptr;
  uint16_t* dl8 = nullptr;       // bf16 d(loss)/d(logit), NHWC padded to 8 channels
  void* wimg_dgrad = nullptr;
  void* a_last_bf16 = nullptr;
  void* wimg_out = nullptr;
  int* tc_error = nullptr;
  int* tc_flag_host = nullptr;   // pinned: the *_host entry points read tc_error back with their results
  // general tensor-core convolution engine (tc_gen.cu): every Conv2D / Conv2DTranspose product the specialised kernels
  // above do not cover runs as a table-driven tcgen05 kernel over bf16 plane tensors
#ifndef KCVAE_EMU
  struct GenLayerPlans {
    GenConvPlan *fwd = nullptr, *fwd_split = nullptr, *dgrad = nullptr;
    GenWgradPlan* wgrad = nullptr;
    size_t img_fwd = 0, img_fwd_split = 0, img_dgrad = 0;      // byte offsets into gen_wimg
  };
  std::vector<GenLayerPlans> gen_e, gen_d;      // encoder convs [L]; decoder: Conv2DTranspose layers [L] + output layer [L]
  bool gen_enc = false;        // encoder forward + backward on the engine
  bool gen_dec = false;        // whole decoder on the engine (no specialised tail for this topology)
  bool gen_dec0 = false;       // README-style decoder: specialised tail, the first Conv2DTranspose's backward on the engine
  bool enc_split = true;       // training forward of the encoder with bf16 hi + lo operand pairs (fp32-grade z)
  bool enc_split_live = false; // layout the encoder plane tensors were last written in
  bool pp_live = false;        // a_pp_planar holds hi + lo planes of the Dense output of the last forward
  bool force_split = false;    // kcvae_loss: loss terms are evaluated with the fp32-grade encoder as well
  unsigned char* gen_wimg = nullptr;
  size_t gen_wimg_bytes = 0;
  std::vector<int32_t> gen_table_host;
  int32_t* gen_table = nullptr;
  uint64_t gen_img_version = 0;
  void* x_pl = nullptr;
  void* x27_pl = nullptr;      // X27 patch planes of the input image (training: shifted operand of the first layer's weight gradient)
  std::vector<void*> act_e_pl, g_e_pl, act_d_pl, g_d_pl;
  bool g_d1_planes_only = false;   // the last backward wrote d loss / d act_d[1] only as planes (g_d_pl[1]), not as fp32
  float *gen_partial = nullptr, *gen_partial2 = nullptr;
  // decoder Dense layer on the engine (GEN_DENSE products): W^T and G^T as plane tensors, plans per batch size
  bool gen_dense = false;
  bool dense_f32_skipped = false;   // the last forward left no fp32 copy of the Dense output (debug_activation unpacks the planes)
  void *wT_pl = nullptr, *gT_pl = nullptr;
  uint64_t wT_version = 0;
  struct DensePlans {
    int B = 0, ones_cap = 0;
    GenConvPlan *fwd = nullptr, *fwd_split = nullptr;
    std::vector<GenConvPlan*> wgrad;           // column blocks of the latent dimension (the last one carries the ones column)
    std::vector<int> wgrad_col0;
    std::vector<size_t> off_wgrad;
    GenWgradPlan* dgrad = nullptr;
    unsigned char* img = nullptr;
    size_t off_fwd = 0, off_fwd_split = 0;
  };
  std::vector<DensePlans> dense_plans;
  // encoder Dense (flatten -> enc_dense units) on the engine: forward = a pixel-K product over the flattened activation
  // (K = flat, hi + lo quadrants, bias through a ones pixel); weight and data gradients = GEN_DENSE forward-type products
  bool gen_edense = false, xT_live = false;
  int ed_F = 0, ed_Fp = 0, ed_E = 0;
  void *xT_pl = nullptr, *wE_pl = nullptr;
  uint64_t wE_version = 0;
  struct EncDensePlans {
    int B = 0;
    GenWgradPlan* fwd = nullptr;
    GenConvPlan *wgrad = nullptr, *dgrad = nullptr;
    unsigned char* img = nullptr;
    size_t off_wgrad = 0, off_dgrad = 0;
  };
  std::vector<EncDensePlans> edense_plans;
#endif
  // data parallel
  int rank = 0, world = 1;
  // collectives run on their own stream so they overlap the backward pass (non-emulated build)
  cudaStream_t comm_stream = nullptr;
  // weight-gradient kernels of the CUDA-core layers run beside the data-gradient kernels of the same layer
  cudaStream_t aux_stream = nullptr;
  cudaEvent_t ev_aux_fork = nullptr, ev_aux_join = nullptr;
  float* partial2 = nullptr;       // reduction scratch of the kernels on aux_stream
  bool use_aux = false, aux_dirty = false;
#ifndef KCVAE_EMU
  cudaEvent_t ev_fork = nullptr, ev_sums = nullptr, ev_comm = nullptr;
#endif
#ifndef KCVAE_EMU
  ncclComm_t comm = nullptr;
#endif

  float* wp(int vi) const { return w + vars[vi].off; }
  float* gp(int vi) const { return g + vars[vi].off; }
  int vi_enc_conv(int l) const { return 2 * l; }
  int vi_enc_dense() const { return 2 * L; }
  int vi_head() const { return 2 * L + (enc_dense ? 2 : 0); }
  int vi_dec_dense() const { return vi_head() + 2; }
  int vi_dec_convT(int l) const { return vi_dec_dense() + 2 + 2 * l; }
  int vi_out() const { return vi_dec_dense() + 2 + 2 * L; }
};

namespace {

#define KC_CUDA(h, expr)                                                                  \
  do {                                                                                    \
    cudaError_t e_ = (expr);                                                              \
    if (e_ != cudaSuccess) {                                                              \
      (h)->err = std::string(#expr) + ": " + cudaGetErrorString(e_);                      \
      return KCVAE_ERR_CUDA;                                                              \
    }                                                                                     \
  } while (0)

int fail(kcvae_model* h, int code, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return code;
}

// true when weight image k must be (re)built before use
bool image_stale(kcvae_model* h, int k) {
  if (!h->w_external && h->img_version[k] == h->w_version) return false;
  h->img_version[k] = h->w_version;
  return true;
}

void pad_before(int n_in, int& before) {  // TF SAME, k=3, s=2
  const int out = (n_in + 1) / 2;
  int total = (out - 1) * 2 + 3 - n_in;
  if (total < 0) total = 0;
  before = total / 2;
}

int build_topology(kcvae_model* h) {
  const kcvae_config& c = h->cfg;
  if (c.image_h <= 0 || c.image_w <= 0 || c.image_c <= 0) return fail(nullptr, KCVAE_ERR_INVALID, "image_size must be positive");
  if (c.n_layers < 0 || c.n_layers > KCVAE_MAX_LAYERS) return fail(nullptr, KCVAE_ERR_INVALID, "unsupported number of layers");
  if (c.latent_dimensions <= 0 || c.latent_dimensions > kMaxLatent) return fail(nullptr, KCVAE_ERR_INVALID, "latent_dimensions out of range (1..1024)");
  if (c.decoder_dense_filters <= 0) return fail(nullptr, KCVAE_ERR_INVALID, "decoder_dense_filters must be positive");
  h->L = c.n_layers; h->H = c.image_h; h->W = c.image_w; h->C = c.image_c;
  h->P = (int64_t)h->H * h->W * h->C;
  h->latent = c.latent_dimensions;
  h->enc_dense = c.encoder_dense_filters > 0 ? c.encoder_dense_filters : 0;
  h->eh = {h->H}; h->ew = {h->W}; h->ec = {h->C};
  for (int l = 0; l < h->L; ++l) {
    if (c.layers[l] <= 0) return fail(nullptr, KCVAE_ERR_INVALID, "layer filters must be positive");
    h->eh.push_back((h->eh.back() + 1) / 2);
    h->ew.push_back((h->ew.back() + 1) / 2);
    h->ec.push_back(c.layers[l]);
  }
  h->flat = h->eh.back() * h->ew.back() * h->ec.back();
  // decoder: int(float(H) / 2^L) (src/abstract_cvae.py:60-61)
  const int h0 = (int)((double)h->H / (double)(1LL << h->L));
  const int w0 = (int)((double)h->W / (double)(1LL << h->L));
  if (h0 == 0) return fail(nullptr, KCVAE_ERR_COLLAPSE, "Error: Build Decoder: Width Collapse: Too many layers, check configuration file: " + std::to_string(h->H) + " -> 0: " + std::to_string(h->L) + " Layers");
  if (w0 == 0) return fail(nullptr, KCVAE_ERR_COLLAPSE, "Error: Build Decoder: Height Collapse: Too many layers, check configuration file: " + std::to_string(h->W) + " -> 0: " + std::to_string(h->L) + " Layers");
  h->dh = {h0}; h->dw = {w0}; h->dc = {c.decoder_dense_filters};
  for (int l = 0; l < h->L; ++l) {
    h->dh.push_back(h->dh.back() * 2);
    h->dw.push_back(h->dw.back() * 2);
    h->dc.push_back(c.layers[h->L - 1 - l]);
  }
  h->dec_units = h0 * w0 * c.decoder_dense_filters;
  // variables
  auto add = [&](std::initializer_list<int64_t> dims) {
    Var v{}; v.rank = (int)dims.size(); v.n = 1; int i = 0;
    for (auto d : dims) { v.dims[i++] = d; v.n *= d; }
    v.off = h->nparams;
    h->nparams += (v.n + 3) / 4 * 4;  // keep every variable 16-byte aligned in the flat vector
    h->vars.push_back(v);
  };
  for (int l = 0; l < h->L; ++l) { add({3, 3, h->ec[l], h->ec[l + 1]}); add({h->ec[l + 1]}); }
  int k = h->flat;
  if (h->enc_dense) { add({k, h->enc_dense}); add({h->enc_dense}); k = h->enc_dense; }
  add({k, 2 * h->latent}); add({2 * h->latent});
  add({h->latent, h->dec_units}); add({h->dec_units});
  for (int l = 0; l < h->L; ++l) { add({3, 3, h->dc[l + 1], h->dc[l]}); add({h->dc[l + 1]}); }
  add({3, 3, h->C, h->dc[h->L]}); add({h->C});
  return KCVAE_OK;
}

#ifndef KCVAE_EMU
// ------------------------------------------------------------------ general convolution engine glue (tc_gen.cu)
int kc16(int C) { return (C + 15) / 16 * 2; }          // 8-channel chunks of a channel count padded to 16
GenPlanes pl_make(void* base, int layout, int KC, int split, int H, int W) {
  GenPlanes p{};
  p.base = base; p.layout = layout; p.KC = KC; p.split = split; p.H = H; p.W = W;
  return p;
}
GenPlanes pl_x(const kcvae_model* h, int split) {        // the input image, 2x2 space-to-depth
  return h->C == 3 ? pl_make(h->x_pl, GEN_X3, 1, split, h->H / 2, h->W / 2) : pl_make(h->x_pl, GEN_S2D, kc16(h->C), split, h->H / 2, h->W / 2);
}
GenPlanes pl_act_e(const kcvae_model* h, int l, int split) {   // encoder activation l in 1..L-1, stored space-to-depth
  return pl_make(h->act_e_pl[l], GEN_S2D, kc16(h->ec[l]), split, h->eh[l] / 2, h->ew[l] / 2);
}
GenPlanes pl_g_e(const kcvae_model* h, int l) {                // encoder gradient l in 1..L
  return pl_make(h->g_e_pl[l], GEN_PLAIN, kc16(h->ec[l]), 0, h->eh[l], h->ew[l]);
}
int kc_act_d(const kcvae_model* h, int l) { return l == 0 ? h->dc[0] / 8 : kc16(h->dc[l]); }   // the Dense writes exactly dc[0] / 8 chunks
GenPlanes pl_act_d(const kcvae_model* h, int l, int split) {   // decoder activation l in 0..L
  void* base = (l == 0 && h->gen_dec0) ? h->a_pp_planar : h->act_d_pl[l];
  return pl_make(base, GEN_PLAIN, kc_act_d(h, l), split, h->dh[l], h->dw[l]);
}
GenPlanes pl_g_d(const kcvae_model* h, int l) {                // decoder gradient l in 1..L, stored space-to-depth
  return pl_make(h->g_d_pl[l], GEN_S2D, kc16(h->dc[l]), 0, h->dh[l] / 2, h->dw[l] / 2);
}

void gen_free_plans(std::vector<kcvae_model::GenLayerPlans>& v) {
  for (auto& g : v) {
    gen_conv_plan_free(g.fwd); gen_conv_plan_free(g.fwd_split); gen_conv_plan_free(g.dgrad); gen_wgrad_plan_free(g.wgrad);
  }
  v.clear();
}

// appends the plan's gather table (indices shifted to the flat parameter vector) and returns the image's byte offset
size_t gen_add_image(kcvae_model* h, const GenConvPlan* p, int vi) {
  size_t n = 0;
  const int32_t* t = gen_conv_table(p, &n);
  const size_t off = h->gen_table_host.size() * 2;
  const int32_t base = (int32_t)h->vars[vi].off;
  for (size_t i = 0; i < n; ++i) {
    const int32_t e = t[i];
    h->gen_table_host.push_back(e < 0 ? -1 : (((e & (GEN_LO_FLAG - 1)) + base) | (e & GEN_LO_FLAG)));
  }
  while (h->gen_table_host.size() % 8) h->gen_table_host.push_back(-1);     // keep every image 16-byte aligned
  return off;
}

// Chooses which layers run on the general engine and builds their plans (shape-based kernel selection, once per handle).
#define KC_TRY_SETUP(expr) do { int rc_ = (expr); if (rc_ != KCVAE_OK) return rc_; } while (0)
int gen_alloc(kcvae_model* h, void** p, size_t units);
int gen_setup(kcvae_model* h) {
  const int L = h->L;
  const char* off = std::getenv("KCVAE_GEN");               // 0 = specialised / CUDA-core kernels only (development switch)
  if (L == 0 || (off && off[0] == '0')) return KCVAE_OK;
  const char* es = std::getenv("KCVAE_ENC_SPLIT");          // 0 = plain bf16 operands in the training forward of the encoder
  h->enc_split = !(es && es[0] == '0');
  const char* why = "";
  // ---- encoder: every Conv2D on the engine, or none
  bool ok = true;
  for (int l = 0; l < L; ++l) ok = ok && h->eh[l] % 2 == 0 && h->ew[l] % 2 == 0;
  std::vector<kcvae_model::GenLayerPlans> E(L);
  for (int l = 0; l < L && ok; ++l) {
    const bool x3 = l == 0 && h->C == 3;
    GenConvSpec f{};
    f.kind = GEN_CONV_S2; f.in_layout = x3 ? GEN_X3 : GEN_S2D; f.Ck = h->ec[l]; f.Cn = h->ec[l + 1]; f.KCk = x3 ? 1 : kc16(h->ec[l]);
    f.w_mode = 0; f.Hg = h->eh[l + 1]; f.Wg = h->ew[l + 1];
    f.split = 0; E[l].fwd = gen_conv_plan_create(f, &why);
    f.split = 1; E[l].fwd_split = gen_conv_plan_create(f, &why);
    if (l > 0) {
      GenConvSpec d{};
      d.kind = GEN_CONVT_S2; d.in_layout = GEN_PLAIN; d.Ck = h->ec[l + 1]; d.Cn = h->ec[l]; d.KCk = kc16(h->ec[l + 1]); d.w_mode = 1;
      d.Hg = h->eh[l + 1]; d.Wg = h->ew[l + 1];
      E[l].dgrad = gen_conv_plan_create(d, &why);
    }
    GenWgradSpec w{};
    w.kind = GEN_CONV_S2; w.s_layout = x3 ? GEN_X27 : f.in_layout; w.s_KC = f.KCk; w.u_layout = GEN_PLAIN; w.u_KC = kc16(h->ec[l + 1]);
    w.Cs = h->ec[l]; w.Cu = h->ec[l + 1]; w.w_mode = 0; w.Hg = h->eh[l + 1]; w.Wg = h->ew[l + 1];
    E[l].wgrad = gen_wgrad_plan_create(w, &why);
    ok = E[l].fwd && E[l].fwd_split && (l == 0 || E[l].dgrad) && E[l].wgrad;
  }
  if (ok) { h->gen_e = E; h->gen_enc = true; } else gen_free_plans(E);
  // ---- decoder
  const bool special_tail = h->use_tc_out;
  const bool dense_planar = h->dc[0] % 8 == 0 && h->dec_units >= 64 && h->latent <= 4096;
  std::vector<kcvae_model::GenLayerPlans> D(L + 1);
  auto convT_plans = [&](int l, bool with_fwd) {
    const int kin = kc_act_d(h, l);
    if (with_fwd) {
      GenConvSpec f{};
      f.kind = GEN_CONVT_S2; f.in_layout = GEN_PLAIN; f.Ck = h->dc[l]; f.Cn = h->dc[l + 1]; f.KCk = kin; f.w_mode = 1;
      f.Hg = h->dh[l]; f.Wg = h->dw[l];
      D[l].fwd = gen_conv_plan_create(f, &why);
    }
    GenConvSpec d{};
    d.kind = GEN_CONV_S2; d.in_layout = GEN_S2D; d.Ck = h->dc[l + 1]; d.Cn = h->dc[l]; d.KCk = kc16(h->dc[l + 1]); d.w_mode = 0;
    d.Hg = h->dh[l]; d.Wg = h->dw[l];
    D[l].dgrad = gen_conv_plan_create(d, &why);
    GenWgradSpec w{};
    w.kind = GEN_CONVT_S2; w.s_layout = GEN_PLAIN; w.s_KC = kin; w.u_layout = GEN_S2D; w.u_KC = kc16(h->dc[l + 1]);
    w.Cs = h->dc[l]; w.Cu = h->dc[l + 1]; w.w_mode = 1; w.Hg = h->dh[l]; w.Wg = h->dw[l];
    D[l].wgrad = gen_wgrad_plan_create(w, &why);
    return (!with_fwd || D[l].fwd) && D[l].dgrad && D[l].wgrad;
  };
  if (!special_tail && dense_planar && h->C <= 8 && h->dh[L] == h->H && h->dw[L] == h->W) {
    bool dok = true;
    for (int l = 0; l < L && dok; ++l) dok = convT_plans(l, true);
    if (dok) {
      GenConvSpec f{};
      f.kind = GEN_CONV_S1; f.in_layout = GEN_PLAIN; f.Ck = h->dc[L]; f.Cn = h->C; f.KCk = kc_act_d(h, L); f.w_mode = 1; f.flip = 1;
      f.Hg = h->H; f.Wg = h->W;
      D[L].fwd = gen_conv_plan_create(f, &why);
      GenConvSpec d{};
      d.kind = GEN_CONV_S1; d.in_layout = GEN_PLAIN; d.Ck = h->C; d.Cn = h->dc[L]; d.KCk = 1; d.w_mode = 0; d.flip = 0;
      d.Hg = h->H; d.Wg = h->W;
      D[L].dgrad = gen_conv_plan_create(d, &why);
      GenWgradSpec w{};
      w.kind = GEN_CONV_S1; w.flip = 1; w.s_layout = GEN_PLAIN; w.s_KC = kc_act_d(h, L); w.u_layout = GEN_PLAIN; w.u_KC = 1;
      w.Cs = h->dc[L]; w.Cu = h->C; w.w_mode = 1; w.Hg = h->H; w.Wg = h->W;
      D[L].wgrad = gen_wgrad_plan_create(w, &why);
      dok = D[L].fwd && D[L].dgrad && D[L].wgrad;
    }
    if (dok) { h->gen_d = D; h->gen_dec = true; } else gen_free_plans(D);
  } else if (special_tail && h->use_tc_convT && h->use_tc_convT_few && h->use_tc_convT_few_train && h->use_tc_convT_bwd && L == 2 &&
             dense_planar) {
    // README-style decoder: the specialised tail kernels stay; the backward of the first Conv2DTranspose joins the engine
    if (convT_plans(0, false)) { h->gen_d = D; h->gen_dec0 = true; } else gen_free_plans(D);
  }
  {
    const char* gd = std::getenv("KCVAE_GEN_DENSE");        // 0 = decoder Dense on the fp32 CUDA-core kernels
    const int vi = h->vi_dec_dense();
    h->gen_dense = (h->gen_dec || h->gen_dec0) && dense_planar && h->dec_units % 32 == 0 && !(gd && gd[0] == '0') &&
                   h->vars[vi + 1].off == h->vars[vi].off + (int64_t)h->latent * h->dec_units;     // dW and db contiguous: one product writes both
    if (h->gen_dense) KC_TRY_SETUP(gen_alloc(h, &h->wT_pl, (size_t)2 * ((h->latent + 7) / 8) * h->dec_units));
  }
  {
    const char* gd = std::getenv("KCVAE_GEN_DENSE");
    // the Dense right behind Flatten: K = flat (thousands), few outputs; the 16 -> 64 head behind it stays a CUDA-core GEMM
    // measured at 256 frames (profiles/r03_b_*): forward 0.061 ms and backward 0.12 ms through the engine against 0.042 / 0.09 ms for
    // the split-K CUDA-core GEMMs (K = 21,000, 16 outputs: no reuse for a tensor core to exploit) - implemented, verified,
    // and OFF unless KCVAE_GEN_EDENSE=1
    const char* ge = std::getenv("KCVAE_GEN_EDENSE");
    h->gen_edense = h->gen_enc && h->enc_dense > 0 && h->enc_dense <= 128 && h->flat >= 1024 && !(gd && gd[0] == '0') && ge && ge[0] == '1';
    if (h->gen_edense) {
      h->ed_F = h->flat; h->ed_E = h->enc_dense; h->ed_Fp = (h->flat + 1 + 31) / 32 * 32;
      KC_TRY_SETUP(gen_alloc(h, &h->wE_pl, (size_t)2 * ((h->ed_E + 7) / 8) * h->ed_Fp));
    }
  }
  if (!h->gen_enc && !h->gen_dec && !h->gen_dec0) return KCVAE_OK;
  // ---- weight images: one gather table for all plans
  for (int l = 0; l < L && h->gen_enc; ++l) {
    const int vi = h->vi_enc_conv(l);
    h->gen_e[l].img_fwd = gen_add_image(h, h->gen_e[l].fwd, vi);
    h->gen_e[l].img_fwd_split = gen_add_image(h, h->gen_e[l].fwd_split, vi);
    if (h->gen_e[l].dgrad) h->gen_e[l].img_dgrad = gen_add_image(h, h->gen_e[l].dgrad, vi);
  }
  for (int l = 0; l <= L && (h->gen_dec || h->gen_dec0); ++l) {
    if (l >= (int)h->gen_d.size()) break;
    const int vi = l < L ? h->vi_dec_convT(l) : h->vi_out();
    if (h->gen_d[l].fwd) h->gen_d[l].img_fwd = gen_add_image(h, h->gen_d[l].fwd, vi);
    if (h->gen_d[l].dgrad) h->gen_d[l].img_dgrad = gen_add_image(h, h->gen_d[l].dgrad, vi);
  }
  h->gen_wimg_bytes = h->gen_table_host.size() * 2;
  KC_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->gen_wimg), h->gen_wimg_bytes + 16));
  KC_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->gen_table), h->gen_table_host.size() * sizeof(int32_t) + 16));
  KC_CUDA(h, cudaMemcpy(h->gen_table, h->gen_table_host.data(), h->gen_table_host.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  KC_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->gen_partial), gen_wgrad_partial_floats(nullptr) * sizeof(float)));
  KC_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->gen_partial2), gen_wgrad_partial_floats(nullptr) * sizeof(float)));
  return KCVAE_OK;
}

// all bf16 weight images of the engine in one launch, only after the weights changed
void gen_refresh(kcvae_model* h, cudaStream_t st) {
  if (!h->gen_wimg || (!h->w_external && h->gen_img_version == h->w_version)) return;
  g_tag = "step";
  gen_gather_weights(h->w, h->gen_table, (int64_t)h->gen_table_host.size(), h->gen_wimg, st);
  h->gen_img_version = h->w_version;
}

int gen_alloc(kcvae_model* h, void** p, size_t units) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  KC_CUDA(h, cudaMalloc(p, (units ? units : 1) * 16));
  return KCVAE_OK;
}

// ---- decoder Dense (src/abstract_cvae.py:75-77) on the engine ---------------------------------------------------------
void dense_plans_free(kcvae_model::DensePlans& d) {
  gen_conv_plan_free(d.fwd); gen_conv_plan_free(d.fwd_split);
  for (GenConvPlan* q : d.wgrad) gen_conv_plan_free(q);
  gen_wgrad_plan_free(d.dgrad);
  if (d.img) cudaFree(d.img);
  d = kcvae_model::DensePlans();
}
// plans of the three Dense products for batch size B (the batch is a GEMM dimension here: columns of the forward product,
// K of the weight gradient); nullptr when the engine does not take this size (the CUDA-core kernels run instead)
kcvae_model::DensePlans* dense_plans_get(kcvae_model* h, int B) {
  // measured (profiles/r02_x_*): at 256 frames the three engine products replace 0.40 ms of CUDA-core kernels with 0.25 ms; at
  // 32 frames the Dense layer is bound by its 17 MB weight matrix either way and the extra pack / gather launches cost 0.02 ms
  const char* mb = std::getenv("KCVAE_GEN_DENSE_MIN_BATCH");       // tests lower it so that small batches drive the engine path too
  const int min_batch = mb ? std::atoi(mb) : 64;
  if (!h->gen_dense || B > 1024 || B < min_batch) return nullptr;
  for (auto& d : h->dense_plans) {
    if (d.B == B && d.ones_cap == h->cap_fwd) return &d;
  }
  for (auto& d : h->dense_plans) if (d.B == B) dense_plans_free(d);      // the ones slot moved with the workspace
  const int K = h->latent, N = h->dec_units, NR = N / 32;
  const char* why = "";
  kcvae_model::DensePlans d;
  d.B = B; d.ones_cap = h->cap_fwd;
  GenConvSpec f{};
  f.kind = GEN_DENSE; f.in_layout = GEN_PLAIN; f.Ck = K; f.KCk = (K + 7) / 8; f.Cn = B; f.w_mode = 1; f.Hg = NR; f.Wg = 32;
  bool ok = true;
  if (B <= 256) {            // the batch is the column dimension of the forward product (<= 256 accumulator columns); beyond
    f.split = 0; d.fwd = gen_conv_plan_create(f, &why);       // that the forward stays on the streaming CUDA-core kernel and
    f.split = 1; d.fwd_split = gen_conv_plan_create(f, &why); // only the two backward products run here
    ok = d.fwd && d.fwd_split;
  }
  for (int c0 = 0; c0 < K && ok; c0 += 240) {
    const int nc = std::min(240, K - c0);
    const bool last = c0 + nc == K;
    GenConvSpec w{};
    w.kind = GEN_DENSE; w.in_layout = GEN_PLAIN; w.Ck = B; w.KCk = (B + 7) / 8; w.Cn = nc + (last ? 1 : 0); w.w_mode = 0;
    w.w_stride = K; w.w_col0 = c0; w.ones_col1 = last ? nc + 1 : 0; w.ones_src = h->cap_fwd * K; w.Hg = NR; w.Wg = 32;
    GenConvPlan* q = gen_conv_plan_create(w, &why);
    ok = q != nullptr;
    if (q) { d.wgrad.push_back(q); d.wgrad_col0.push_back(c0); }
  }
  if (ok) {
    GenWgradSpec g{};
    g.kind = GEN_DENSE; g.s_layout = GEN_PLAIN; g.s_KC = (K + 7) / 8; g.u_layout = GEN_PLAIN; g.u_KC = (B + 7) / 8;
    g.Cs = K; g.Cu = B; g.w_mode = 1; g.Hg = NR; g.Wg = 32;
    d.dgrad = gen_wgrad_plan_create(g, &why);
    ok = d.dgrad != nullptr;
  }
  if (ok) {
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    size_t total = 0;
    if (d.fwd) {
      d.off_fwd = total; total += up(gen_conv_weight_image_bytes(d.fwd));
      d.off_fwd_split = total; total += up(gen_conv_weight_image_bytes(d.fwd_split));
    }
    for (GenConvPlan* q : d.wgrad) { d.off_wgrad.push_back(total); total += up(gen_conv_weight_image_bytes(q)); }
    ok = cudaMalloc(reinterpret_cast<void**>(&d.img), total + 256) == cudaSuccess;
  }
  if (!ok) { dense_plans_free(d); return nullptr; }
  h->dense_plans.push_back(d);
  return &h->dense_plans.back();
}

// W^T of the decoder Dense as planes (hi + lo), once per weight version
void dense_refresh_wT(kcvae_model* h, cudaStream_t st) {
  if (!h->w_external && h->wT_version == h->w_version) return;
  gen_pack_rows_T(h->wp(h->vi_dec_dense()), h->latent, h->dec_units, 1, h->wT_pl, st);
  h->wT_version = h->w_version;
}
// relu(z W + b) -> the bf16 image of the Dense output (planes_out, hi + lo when split) and / or fp32 [B][N]
bool gen_dense_forward(kcvae_model* h, const float* z, int B, int split, void* planes_out, float* f32_out, cudaStream_t st) {
  // measured at 256 frames (profiles/r03_b_*): 0.147 ms here (W^T pack, z gather, product with its transposing 2-byte stores)
  // against 0.104 ms for the streaming CUDA-core kernel, which writes the same hi + lo planes: the forward stays there unless
  // asked for; the two backward products (0.12 ms against 0.30 ms) are the ones that run on the engine by default
  const char* fw = std::getenv("KCVAE_GEN_DENSE_FWD");
  if (!(fw && fw[0] == '1')) return false;
  kcvae_model::DensePlans* dp = dense_plans_get(h, B);
  if (!dp || !dp->fwd) return false;
  const int vi = h->vi_dec_dense(), K = h->latent, N = h->dec_units;
  dense_refresh_wT(h, st);
  GenConvPlan* plan = split ? dp->fwd_split : dp->fwd;
  unsigned char* img = dp->img + (split ? dp->off_fwd_split : dp->off_fwd);
  gen_conv_prep_weights(plan, z, img, st);                       // this step's z as the B operand
  GenPlanes in = pl_make(h->wT_pl, GEN_PLAIN, (K + 7) / 8, 1, N / 32, 32);
  GenPlanes outp = pl_make(planes_out, GEN_PLAIN, h->dc[0] / 8, split, h->dh[0], h->dw[0]);
  GenEpilogue e{};
  e.pre = GEN_PRE_BIAS_RELU; e.bias = h->wp(vi + 1);
  e.out = planes_out ? &outp : nullptr; e.out_f32 = f32_out;
  e.dense_n = N; e.dense_ld = N; e.dense_cc = h->dc[0];
  if (gen_conv_run(plan, in, img, e, 1, h->tc_error, "gen_dense", st) != 0) h->tc_failed = true;
  return true;
}

// ---- encoder Dense (src/abstract_cvae.py:41-44) on the engine ---------------------------------------------------------
void edense_plans_free(kcvae_model::EncDensePlans& d) {
  gen_wgrad_plan_free(d.fwd); gen_conv_plan_free(d.wgrad); gen_conv_plan_free(d.dgrad);
  if (d.img) cudaFree(d.img);
  d = kcvae_model::EncDensePlans();
}
kcvae_model::EncDensePlans* edense_plans_get(kcvae_model* h, int B) {
  const char* mb = std::getenv("KCVAE_GEN_DENSE_MIN_BATCH");
  const int min_batch = mb ? std::atoi(mb) : 64;
  if (!h->gen_edense || B > 256 || B < min_batch) return nullptr;
  for (auto& d : h->edense_plans) if (d.B == B) return &d;
  const int F = h->ed_F, Fp = h->ed_Fp, E = h->ed_E, KCb = (B + 7) / 8, KCe = (E + 7) / 8;
  const char* why = "";
  kcvae_model::EncDensePlans d;
  d.B = B;
  GenWgradSpec f{};        // d1[b][e] = sum_i flat[b][i] W[i][e] + bias[e]: S = weight planes, U = transposed activation planes
  f.kind = GEN_DENSE; f.s_layout = GEN_PLAIN; f.s_KC = 2 * KCe; f.u_layout = GEN_PLAIN; f.u_KC = 2 * KCb; f.Cs = E; f.Cu = B; f.w_mode = 1;
  f.Hg = Fp / 32; f.Wg = 32; f.split_dense = 1;
  d.fwd = gen_wgrad_plan_create(f, &why);
  GenConvSpec w{};         // dW[i][e] = sum_b flat[b][i] g[b][e]: rows i, K = frames, columns e
  w.kind = GEN_DENSE; w.in_layout = GEN_PLAIN; w.Ck = B; w.KCk = KCb; w.Cn = E; w.w_mode = 0; w.w_stride = E; w.Hg = Fp / 32; w.Wg = 32;
  d.wgrad = gen_conv_plan_create(w, &why);
  GenConvSpec g{};         // g_flat[b][i] = sum_e g[b][e] W[i][e]: rows i, K = e, columns b
  g.kind = GEN_DENSE; g.in_layout = GEN_PLAIN; g.Ck = E; g.KCk = KCe; g.Cn = B; g.w_mode = 1; g.w_stride = E; g.Hg = Fp / 32; g.Wg = 32;
  d.dgrad = gen_conv_plan_create(g, &why);
  bool ok = d.fwd && d.wgrad && d.dgrad;
  if (ok) {
    auto up = [](size_t v) { return (v + 255) / 256 * 256; };
    d.off_wgrad = 0;
    d.off_dgrad = up(gen_conv_weight_image_bytes(d.wgrad));
    ok = cudaMalloc(reinterpret_cast<void**>(&d.img), d.off_dgrad + up(gen_conv_weight_image_bytes(d.dgrad)) + 256) == cudaSuccess;
  }
  if (!ok) { edense_plans_free(d); return nullptr; }
  h->edense_plans.push_back(d);
  return &h->edense_plans.back();
}
// d1 = flat W + b (linear) through the engine; false = not taken (the CUDA-core GEMM runs)
bool gen_edense_forward(kcvae_model* h, const float* flat, int B, float* out, cudaStream_t st) {
  h->xT_live = false;
  kcvae_model::EncDensePlans* ep = edense_plans_get(h, B);
  if (!ep || !h->xT_pl) return false;
  const int vi = h->vi_enc_dense(), F = h->ed_F, Fp = h->ed_Fp, E = h->ed_E;
  if (h->w_external || h->wE_version != h->w_version) {       // weight planes (hi + lo) with the bias in the ones pixel
    gen_pack_cols(h->wp(vi), h->wp(vi + 1), F, E, 1, h->wE_pl, Fp, F, st);
    h->wE_version = h->w_version;
  }
  gen_pack_rows_T(flat, B, F, 1, h->xT_pl, st, Fp, F);         // [frames / 8 (hi | lo)][flat index][8 frames], ones pixel at index F
  GenPlanes S = pl_make(h->wE_pl, GEN_PLAIN, 2 * ((E + 7) / 8), 0, Fp / 32, 32);
  GenPlanes U = pl_make(h->xT_pl, GEN_PLAIN, 2 * ((B + 7) / 8), 0, Fp / 32, 32);
  if (gen_wgrad_run(ep->fwd, S, U, out, nullptr, h->gen_partial, 1, h->tc_error, "gen_edense", st) != 0) h->tc_failed = true;
  h->xT_live = true;
  return true;
}
#endif

template <typename T>
int dalloc(kcvae_model* h, T** p, size_t n) {
  if (*p) { cudaFree(*p); *p = nullptr; }
  if (n == 0) n = 1;
  KC_CUDA(h, cudaMalloc(reinterpret_cast<void**>(p), n * sizeof(T)));
  return KCVAE_OK;
}
#define KC_TRY(expr) do { int rc_ = (expr); if (rc_ != KCVAE_OK) return rc_; } while (0)

size_t max_partial_floats(const kcvae_model* h, int B) {
  size_t mx = 1;
  auto up = [&](size_t v) { if (v > mx) mx = v; };
  const int L = h->L;
  for (int l = 0; l < L; ++l) {
    up(wgrad_partial_floats(B, h->eh[l + 1], h->ec[l + 1], h->ec[l]));
    up(colsum_partial_floats((int64_t)B * h->eh[l + 1] * h->ew[l + 1], h->ec[l + 1]));
    up(wgrad_partial_floats(B, h->dh[l], h->dc[l], h->dc[l + 1]));
    up(colsum_partial_floats((int64_t)B * h->dh[l + 1] * h->dw[l + 1], h->dc[l + 1]));
  }
  up(wgrad_partial_floats(B, h->dh[L], h->C, h->dc[L]));
  up(colsum_partial_floats((int64_t)B * h->dh[L] * h->dw[L], h->C));
  const int kin = h->enc_dense ? h->enc_dense : h->flat;
  up(gemm_partial_floats(B, h->enc_dense, h->flat));
  up(gemm_partial_floats(B, 2 * h->latent, kin));
  up(gemm_partial_floats(B, h->dec_units, h->latent));
  up(gemm_partial_floats(h->latent, h->dec_units, B));
  up(gemm_partial_floats(B, h->latent, h->dec_units));
  up(dense_wide_partial_floats(B, h->dec_units, h->latent));
  up(gemm_partial_floats(kin, 2 * h->latent, B));
  up(gemm_partial_floats(B, kin, 2 * h->latent));
  up(gemm_partial_floats(h->flat, h->enc_dense, B));
  up(gemm_partial_floats(B, h->flat, h->enc_dense));
  up(colsum_partial_floats(B, 2 * h->latent));
  up(colsum_partial_floats(B, h->enc_dense ? h->enc_dense : 1));
  up(colsum_partial_floats(B, h->dec_units));
  up(score_partial_floats(B, (int64_t)h->H * h->W));
#ifndef KCVAE_EMU
  if (h->use_tc_dgrad) up(tc_out_wgrad_partial_floats(h->dc[L], h->C));
  if (h->use_tc_convT_bwd) up(tc_convT_wgrad_partial_floats(h->dc[L - 1]));
  if (h->use_tc_dgrad) up((size_t)kNumSMs * 16 * 32);
  if (h->use_tc_convT && h->use_tc_out) up(tc_tail_score_partial_floats(B, h->H, h->W));
#endif
  return mx;
}

int ensure_fwd(kcvae_model* h, int B) {
  if (B <= h->cap_fwd) return KCVAE_OK;
  const int L = h->L;
  h->act_e.resize(L + 1, nullptr);
  h->act_d.resize(L + 1, nullptr);
#ifndef KCVAE_EMU
  const bool gen_enc = h->gen_enc, gen_dec = h->gen_dec;
#else
  const bool gen_enc = false, gen_dec = false;
#endif
  // fp32 NHWC activations exist only where a CUDA-core / specialised kernel or a Dense layer reads them
  for (int l = 1; l <= L; ++l)
    if (!gen_enc || l == L) KC_TRY(dalloc(h, &h->act_e[l], (size_t)B * h->eh[l] * h->ew[l] * h->ec[l]));
  for (int l = 0; l <= L; ++l)
    if (!gen_dec) KC_TRY(dalloc(h, &h->act_d[l], (size_t)B * h->dh[l] * h->dw[l] * h->dc[l]));
#ifndef KCVAE_EMU
  if (h->gen_edense) KC_TRY(gen_alloc(h, &h->xT_pl, (size_t)2 * ((B + 7) / 8) * h->ed_Fp));
  if (gen_enc) {
    h->act_e_pl.resize(L + 1, nullptr);
    KC_TRY(gen_alloc(h, &h->x_pl, pl_x(h, 1).units(B)));
    if (h->C == 3) KC_TRY(gen_alloc(h, &h->x27_pl, (size_t)B * 4 * (h->H / 2) * (h->W / 2)));
    for (int l = 1; l < L; ++l) KC_TRY(gen_alloc(h, &h->act_e_pl[l], pl_act_e(h, l, 1).units(B)));
  }
  if (gen_dec) {
    h->act_d_pl.resize(L + 1, nullptr);
    for (int l = 0; l <= L; ++l) KC_TRY(gen_alloc(h, &h->act_d_pl[l], pl_act_d(h, l, 0).units(B)));
  }
#endif
  KC_TRY(dalloc(h, &h->x_stage[0], (size_t)B * h->P));
  KC_TRY(dalloc(h, &h->x_stage[1], (size_t)B * h->P));
  h->x_in = h->x_stage[0];
  h->pend_src[0] = h->pend_src[1] = nullptr;
  KC_TRY(dalloc(h, &h->x_noisy, (size_t)B * h->P));
  KC_TRY(dalloc(h, &h->d1, (size_t)B * (h->enc_dense ? h->enc_dense : 1)));
  KC_TRY(dalloc(h, &h->head, (size_t)B * 2 * h->latent));
  KC_TRY(dalloc(h, &h->z, (size_t)B * h->latent + 4));       // + a 1.0f behind it: the "ones column" source of the Dense bias gradient
  {
    const float one = 1.0f;
    KC_CUDA(h, cudaMemcpy(h->z + (size_t)B * h->latent, &one, sizeof(float), cudaMemcpyHostToDevice));
  }
  KC_TRY(dalloc(h, &h->mean, (size_t)B * h->latent));
  KC_TRY(dalloc(h, &h->logvar, (size_t)B * h->latent));
  KC_TRY(dalloc(h, &h->eps_buf, (size_t)B * h->latent));
  KC_TRY(dalloc(h, &h->xhat, (size_t)B * h->dh[L] * h->dw[L] * h->C));
  KC_TRY(dalloc(h, &h->err_buf, (size_t)B * h->H * h->W));
  KC_TRY(dalloc(h, &h->score_buf, (size_t)B));
#ifndef KCVAE_EMU
  if (h->use_tc_out) {
    unsigned short* tmp = reinterpret_cast<unsigned short*>(h->a_last_bf16);
    KC_TRY(dalloc(h, &tmp, (size_t)B * h->dh[L] * h->dw[L] * h->dc[L]));
    h->a_last_bf16 = tmp;
    KC_TRY(dalloc(h, &h->relu_bits, (size_t)B * h->dh[L] * h->dw[L]));
    if (h->use_tc_convT) {
      unsigned short* t8 = reinterpret_cast<unsigned short*>(h->a_prev8);
      KC_TRY(dalloc(h, &t8, (size_t)B * h->dh[L - 1] * h->dw[L - 1] * 8));
      h->a_prev8 = t8;
      if (h->use_tc_convT_few) {
        unsigned short* tp = reinterpret_cast<unsigned short*>(h->a_pp_planar);
        KC_TRY(dalloc(h, &tp, (size_t)2 * B * h->dh[L - 2] * h->dw[L - 2] * h->dc[L - 2]));   // hi + lo planes
        h->a_pp_planar = tp;
      }
    }
  }
#endif
  h->partial_floats = max_partial_floats(h, B);
  KC_TRY(dalloc(h, &h->partial, h->partial_floats));
#ifndef KCVAE_EMU
  {
    const char* e = std::getenv("KCVAE_AUX_STREAM");     // 0 = every backward kernel on the caller's stream
    h->use_aux = !(e && e[0] == '0');
    if (h->use_aux) {
      KC_TRY(dalloc(h, &h->partial2, h->partial_floats));
      if (!h->aux_stream) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);      // side stream at the lowest priority: it only fills gaps
        KC_CUDA(h, cudaStreamCreateWithPriority(&h->aux_stream, cudaStreamNonBlocking, prio_lo));
        KC_CUDA(h, cudaEventCreateWithFlags(&h->ev_aux_fork, cudaEventDisableTiming));
        KC_CUDA(h, cudaEventCreateWithFlags(&h->ev_aux_join, cudaEventDisableTiming));
      }
    }
  }
#endif
  h->cap_fwd = B;
  h->cap_bwd = 0;  // partial buffer was sized for this B; backward buffers follow
  return KCVAE_OK;
}

int ensure_bwd(kcvae_model* h, int B) {
  KC_TRY(ensure_fwd(h, B));
  if (B <= h->cap_bwd) return KCVAE_OK;
  const int L = h->L;
  const int Bc = h->cap_fwd;
  h->g_act_e.resize(L + 1, nullptr);
  h->g_act_d.resize(L + 1, nullptr);
#ifndef KCVAE_EMU
  const bool gen_enc = h->gen_enc, gen_dec = h->gen_dec;
#else
  const bool gen_enc = false, gen_dec = false;
#endif
  for (int l = 1; l <= L; ++l)
    if (!gen_enc || l == L) KC_TRY(dalloc(h, &h->g_act_e[l], (size_t)Bc * h->eh[l] * h->ew[l] * h->ec[l]));
  for (int l = 0; l <= L; ++l)
    if (!gen_dec || l == 0) KC_TRY(dalloc(h, &h->g_act_d[l], (size_t)Bc * h->dh[l] * h->dw[l] * h->dc[l]));
#ifndef KCVAE_EMU
  if (gen_enc) {
    h->g_e_pl.resize(L + 1, nullptr);
    for (int l = 1; l <= L; ++l) KC_TRY(gen_alloc(h, &h->g_e_pl[l], pl_g_e(h, l).units(Bc)));
  }
  if (h->gen_dense) KC_TRY(gen_alloc(h, &h->gT_pl, (size_t)((Bc + 7) / 8) * h->dec_units));
  if (gen_dec || h->gen_dec0) {
    h->g_d_pl.resize(L + 1, nullptr);
    for (int l = 1; l <= (gen_dec ? L : 1); ++l) {
      KC_TRY(gen_alloc(h, &h->g_d_pl[l], pl_g_d(h, l).units(Bc)));
      // planes of channels a producer never writes (8..15 of a 5-channel gradient) must read as zeros
      KC_CUDA(h, cudaMemset(h->g_d_pl[l], 0, pl_g_d(h, l).units(Bc) * 16));
    }
  }
#endif
  KC_TRY(dalloc(h, &h->dlogit, (size_t)Bc * h->P));
  KC_TRY(dalloc(h, &h->g_z, (size_t)Bc * h->latent));
  KC_TRY(dalloc(h, &h->dhead, (size_t)Bc * 2 * h->latent));
  KC_TRY(dalloc(h, &h->g_d1, (size_t)Bc * (h->enc_dense ? h->enc_dense : 1)));
  if (h->use_tc_convT_bwd) {
    unsigned short* gs = reinterpret_cast<unsigned short*>(h->g_s2d);
    KC_TRY(dalloc(h, &gs, (size_t)Bc * h->H * h->W * h->dc[L]));
    h->g_s2d = gs;
  }
  if (h->use_tc_dgrad || gen_dec) {
    KC_TRY(dalloc(h, &h->dl8, (size_t)Bc * h->H * h->W * 8));
    KC_CUDA(h, cudaMemset(h->dl8, 0, (size_t)Bc * h->H * h->W * 8 * sizeof(uint16_t)));   // channel padding stays zero
  }
  h->cap_bwd = Bc;
  return KCVAE_OK;
}

int check_batch(kcvae_model* h, int B) {
  if (!h) return KCVAE_ERR_INVALID;
  if (B <= 0) return fail(h, KCVAE_ERR_INVALID, "batch must be positive");
  return KCVAE_OK;
}
int check_recon_shape(kcvae_model* h) {
  if (h->dh[h->L] != h->H || h->dw[h->L] != h->W)
    return fail(h, KCVAE_ERR_INVALID, "decoder output " + std::to_string(h->dh[h->L]) + "x" + std::to_string(h->dw[h->L]) +
                " does not match image " + std::to_string(h->H) + "x" + std::to_string(h->W) +
                ": image_size must be divisible by 2^len(layers) (x - x_hat does not broadcast)");
  return KCVAE_OK;
}
int post(kcvae_model* h) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) { h->err = std::string("kernel launch: ") + cudaGetErrorString(e); return KCVAE_ERR_CUDA; }
  return KCVAE_OK;
}

// A tensor-core launcher that could not run (cuTensorMapEncodeTiled) invalidates the call: there is no fallback kernel.
int tc_check(kcvae_model* h) {
  if (!h->tc_failed) return KCVAE_OK;
  h->tc_failed = false;
  return fail(h, KCVAE_ERR_CUDA, "tensor-core path: cuTensorMapEncodeTiled failed; use precision fp32");
}
// Host entry points synchronise anyway: read the device flag the bounded tcgen05 barrier waits raise and refuse the
// result when it is set (an expired wait must not feed garbage gradients to Adam silently).
int tc_flag_check(kcvae_model* h, cudaStream_t st) {   // enqueue the flag read-back, synchronise, test
  if (h->tc_error && h->tc_flag_host) KC_CUDA(h, cudaMemcpyAsync(h->tc_flag_host, h->tc_error, sizeof(int), cudaMemcpyDeviceToHost, st));
  KC_CUDA(h, cudaStreamSynchronize(st));
  if (h->tc_error && h->tc_flag_host && *h->tc_flag_host)
    return fail(h, KCVAE_ERR_CUDA, "tcgen05 pipeline: bounded mbarrier wait expired (results of this call are invalid)");
  return KCVAE_OK;
}

// ------------------------------------------------------------------------------ forward
#ifndef KCVAE_EMU
// encoder convolutions on the general engine: x -> 2x2 space-to-depth bf16 planes (hi + lo when `split`), every Conv2D s2
// as a stride-1 product over them; the last activation leaves as fp32 NHWC for the Dense layers
void gen_run_encoder_convs(kcvae_model* h, const float* x, int B, int split, cudaStream_t st, int train) {
  const int L = h->L;
  gen_refresh(h, st);
  h->enc_split_live = split != 0;
  GenPlanes in = pl_x(h, split);
  g_tag = "enc.pack";
  if (h->C == 3) gen_pack_x3(x, B, h->H, h->W, split, h->x_pl, st, train ? h->x27_pl : nullptr);
  else gen_pack_nhwc(x, B, h->H, h->W, h->C, in, st);
  for (int l = 0; l < L; ++l) {
    const auto& g = h->gen_e[l];
    GenEpilogue e{};
    e.pre = GEN_PRE_BIAS_RELU; e.bias = h->wp(h->vi_enc_conv(l) + 1);
    GenPlanes out{};
    if (l + 1 < L) { out = pl_act_e(h, l + 1, split); e.out = &out; }
    else e.out_f32 = h->act_e[L];
    g_tag = l == 0 ? "enc.conv0.fwd" : (l == 1 ? "enc.conv1.fwd" : "enc.convN.fwd");
    if (gen_conv_run(split ? g.fwd_split : g.fwd, in, h->gen_wimg + (split ? g.img_fwd_split : g.img_fwd), e, B, h->tc_error,
                     "gen_conv", st) != 0) h->tc_failed = true;
    in = out;
  }
}
#endif

void run_encoder(kcvae_model* h, const float* x, int B, cudaStream_t st, int split = 0) {
  const float* in = x;
#ifndef KCVAE_EMU
  if (h->gen_enc) {
    gen_run_encoder_convs(h, x, B, split && h->enc_split, st, split);
    in = h->act_e[h->L];
  } else
#endif
  for (int l = 0; l < h->L; ++l) {
    ConvArgs a{};
    a.in = in; a.w = h->wp(h->vi_enc_conv(l)); a.bias = h->wp(h->vi_enc_conv(l) + 1); a.out = h->act_e[l + 1];
    a.B = B; a.Hi = h->eh[l]; a.Wi = h->ew[l]; a.Ci = h->ec[l];
    a.Ho = h->eh[l + 1]; a.Wo = h->ew[l + 1]; a.Co = h->ec[l + 1];
    a.w_sci = a.Co; a.w_sco = 1;  // HWIO
    pad_before(a.Hi, a.pad_t); pad_before(a.Wi, a.pad_l);
    g_tag = l == 0 ? "enc.conv0.fwd" : (l == 1 ? "enc.conv1.fwd" : "enc.convN.fwd");
    conv_forward(CONV_S2, EPI_BIAS_RELU, a, st);
    in = h->act_e[l + 1];
  }
  const float* flat = in;
  int k = h->flat;
  if (h->enc_dense) {
    GemmArgs ga{};
    ga.A = flat; ga.a_sm = k; ga.a_sk = 1;
    ga.Bm = h->wp(h->vi_enc_dense()); ga.b_sk = h->enc_dense; ga.b_sn = 1;
    ga.C = h->d1; ga.bias = h->wp(h->vi_enc_dense() + 1);
    ga.M = B; ga.N = h->enc_dense; ga.K = k; ga.partial = h->partial;
    g_tag = "enc.dense.fwd";
#ifndef KCVAE_EMU
    if (!gen_edense_forward(h, flat, B, h->d1, st))
#endif
    gemm(ga, st);
    flat = h->d1; k = h->enc_dense;
  }
  GemmArgs ga{};
  ga.A = flat; ga.a_sm = k; ga.a_sk = 1;
  ga.Bm = h->wp(h->vi_head()); ga.b_sk = 2 * h->latent; ga.b_sn = 1;
  ga.C = h->head; ga.bias = h->wp(h->vi_head() + 1);
  ga.M = B; ga.N = 2 * h->latent; ga.K = k; ga.partial = h->partial;
  g_tag = "enc.head.fwd";
  gemm(ga, st);
}

// true when the two last decoder layers can run as the single fused tensor-core kernel
bool tail_fusable(const kcvae_model* h, int B) {
#ifndef KCVAE_EMU
  const int L = h->L;
  return L >= 1 && h->use_tc_convT && h->use_tc_out && h->dc[L] == 32 &&
         tc_tail_fused_supported(h->dc[L - 1], h->dc[L], h->C, h->dh[L], h->dw[L]) &&
         h->partial_floats >= tc_tail_score_partial_floats(B, h->dh[L], h->dw[L]);
#else
  (void)h; (void)B;
  return false;
#endif
}

// What the fused decoder tail may produce besides / instead of x_hat (scoring, do_anomaly_detection.py:62,88)
struct TailOut {
  const float* x = nullptr;      // frames the reconstruction is compared with (needed for err / score)
  float* err = nullptr;          // [B,H,W] or nullptr
  float* score = nullptr;        // [B] or nullptr
  float* err_minmax = nullptr;   // [B,2] or nullptr
  bool done = false;             // set when the fused kernel produced them
};

// keep_last: the last 32-channel activation must exist in HBM afterwards (training: the backward reads it).
// Inference entry points pass false: with the tensor-core tail the two last layers then run as ONE kernel and
// that activation only ever exists as a shared-memory tile.  out may be nullptr when only the score is wanted.
void run_decoder(kcvae_model* h, const float* z, int B, int apply_sigmoid, float* out, cudaStream_t st,
                 bool keep_last = true, TailOut* tail = nullptr) {
  const int L = h->L;
  bool last_is_bf16 = false;   // the last activation was produced directly in bf16 by tc_convT_fwd
  bool prev8_ready = false;    // a_prev8 was written by the 32 -> few tensor-core layer (no pack pass needed)
  h->relu_bits_valid = false;
  GemmArgs ga{};
  ga.A = z; ga.a_sm = h->latent; ga.a_sk = 1;
  ga.Bm = h->wp(h->vi_dec_dense()); ga.b_sk = h->dec_units; ga.b_sn = 1;
  ga.C = h->act_d[0]; ga.bias = h->wp(h->vi_dec_dense() + 1); ga.relu = 1;
  ga.M = B; ga.N = h->dec_units; ga.K = h->latent; ga.partial = h->partial;
  g_tag = "dec.dense.fwd";
#ifndef KCVAE_EMU
  if (h->gen_dec) {
    // whole decoder on the general engine: Dense -> bf16 planes, every Conv2DTranspose and the output layer as tcgen05 products
    gen_refresh(h, st);
    if (!gen_dense_forward(h, z, B, 0, h->act_d_pl[0], nullptr, st))
      dense_wide_forward(z, ga.Bm, ga.bias, nullptr, B, h->dec_units, h->latent, 1, st, h->act_d_pl[0], h->dc[0], 0);
    for (int l = 0; l <= L; ++l) {
      const auto& g = h->gen_d[l];
      GenPlanes in = pl_act_d(h, l, 0), outp{};
      GenEpilogue e{};
      if (l < L) {
        e.pre = GEN_PRE_BIAS_RELU; e.bias = h->wp(h->vi_dec_convT(l) + 1);
        outp = pl_act_d(h, l + 1, 0); e.out = &outp;
        g_tag = l == L - 1 ? "dec.convT_last.fwd" : (l == L - 2 ? "dec.convT.fwd" : "dec.convT_early.fwd");
      } else {
        e.pre = apply_sigmoid ? GEN_PRE_BIAS_SIGMOID : GEN_PRE_BIAS; e.bias = h->wp(h->vi_out() + 1);
        e.out_f32 = out;
        g_tag = "dec.out.fwd";
      }
      if (gen_conv_run(g.fwd, in, h->gen_wimg + g.img_fwd, e, B, h->tc_error, "gen_conv", st) != 0) h->tc_failed = true;
    }
    return;
  }
#endif
  // the 32 -> few tensor-core Conv2DTranspose reads the Dense output as chunk-planar bf16: when the Dense is the layer right
  // before it, it writes that copy itself (and, with nothing else reading the fp32 activation, only that copy)
  bool few_tc = false, pp_ready = false;
#ifndef KCVAE_EMU
  // training (keep_last): only with the hi + lo operand pairs, which the Dense layer right before has to write
  const bool wide = dense_wide_ok(z, ga.Bm, ga.C, ga.bias, B, h->dec_units, h->latent);
  const bool split = keep_last;
  few_tc = L >= 2 && h->use_tc_convT_few && h->use_tc_convT && h->use_tc_out &&
           (!keep_last || (h->use_tc_convT_few_train && wide && L == 2 && h->dc[0] % 8 == 0));
#else
  const bool wide = dense_wide_ok(z, ga.Bm, ga.C, ga.bias, B, h->dec_units, h->latent);
  const bool split = false;
#endif
  if (wide) {
    pp_ready = few_tc && L == 2 && h->dc[0] % 8 == 0;
#ifndef KCVAE_EMU
    // (training with the engine's backward behind it: the hi + lo planes are the only copy of the Dense output anybody reads)
    h->dense_f32_skipped = false;
    if (pp_ready && gen_dense_forward(h, z, B, split ? 1 : 0, h->a_pp_planar,
                                      (!keep_last || (h->gen_dec0 && split)) ? nullptr : ga.C, st)) {
      h->dense_f32_skipped = keep_last && h->gen_dec0 && split;
    } else
#endif
    dense_wide_forward(z, ga.Bm, ga.bias, (pp_ready && !keep_last) ? nullptr : ga.C, B, h->dec_units, h->latent, 1, st,
                       pp_ready ? h->a_pp_planar : nullptr, h->dc[0], split ? 1 : 0);
#ifndef KCVAE_EMU
    h->pp_live = pp_ready && split;
#endif
  } else {
    gemm(ga, st);
  }
  for (int l = 0; l < L; ++l) {
    ConvArgs a{};
    a.in = h->act_d[l]; a.w = h->wp(h->vi_dec_convT(l)); a.bias = h->wp(h->vi_dec_convT(l) + 1); a.out = h->act_d[l + 1];
    a.B = B; a.Hi = h->dh[l]; a.Wi = h->dw[l]; a.Ci = h->dc[l];
    a.Ho = h->dh[l + 1]; a.Wo = h->dw[l + 1]; a.Co = h->dc[l + 1];
    a.w_sci = 1; a.w_sco = a.Ci;  // [kh,kw,out,in]
    g_tag = l == L - 1 ? "dec.convT_last.fwd" : (l == L - 2 ? "dec.convT.fwd" : "dec.convT_early.fwd");
#ifndef KCVAE_EMU
    if (l == L - 2 && few_tc) {
      // 32 -> few Conv2DTranspose on tcgen05: bf16 chunk-planar copy of the input, output straight into the 8-channel
      // bf16 units the next tensor-core layer reads (+ the fp32 activation when the backward will need it)
      if (!pp_ready) cast_f32_to_bf16_planar(a.in, h->a_pp_planar, B, (int64_t)a.Hi * a.Wi, a.Ci, st);
      if (image_stale(h, 4)) tc_prep_convT_few_weights(a.w, a.Co, a.Ci, h->wimg_convT_few, st);
      if (tc_convT_few_fwd(h->a_pp_planar, h->wimg_convT_few, a.bias, h->a_prev8, keep_last ? a.out : nullptr, B, a.Hi, a.Wi, a.Co,
                           (split && pp_ready) ? 1 : 0, h->tc_error, st) == 0) {
        prev8_ready = true;
        continue;
      }
    }
    if (l == L - 1 && h->use_tc_convT && h->use_tc_out) {
      // sub-pixel phase decomposition on tcgen05, bf16 output straight into the buffer the
      // output-layer kernels read (no fp32 copy of the 224x300x32 activation exists in this mode)
      if (!prev8_ready) pack_c8_bf16(a.in, (int64_t)B * a.Hi * a.Wi, a.Ci, h->a_prev8, st);
      if (image_stale(h, 0)) tc_prep_convT_weights(a.w, a.Co, a.Ci, h->wimg_convT, st);
      if ((!keep_last || h->fuse_train_tail) && tail_fusable(h, B)) {
        const int vo = h->vi_out();
        if (image_stale(h, 1)) tc_prep_tail_weights(h->wp(vo), h->C, h->dc[L], h->wimg_out, st);
        g_tag = "dec.tail";
        const bool want_score = tail && tail->x && (tail->err || tail->score);
        if (tc_tail_fused(h->a_prev8, h->wimg_convT, h->wimg_out, a.bias, h->wp(vo + 1), want_score ? tail->x : nullptr, out,
                          keep_last ? h->a_last_bf16 : nullptr, h->relu_bits, want_score ? tail->err : nullptr,
                          want_score ? tail->score : nullptr,
                          want_score ? tail->err_minmax : nullptr, h->partial, B, h->dh[L], h->dw[L], h->C, apply_sigmoid,
                          h->tc_error, st) == 0) {
          if (want_score) tail->done = true;
          h->relu_bits_valid = keep_last && h->relu_bits;
          return;
        }
        g_tag = "dec.convT_last.fwd";
      }
      if (tc_convT_fwd(h->a_prev8, h->wimg_convT, a.bias, h->a_last_bf16, B, a.Hi, a.Wi, h->tc_error, st) == 0) {
        last_is_bf16 = true;
        continue;
      }
    }
#endif
    conv_forward(CONVT_S2, EPI_BIAS_RELU, a, st);
  }
  ConvArgs a{};
  a.in = h->act_d[L]; a.w = h->wp(h->vi_out()); a.bias = h->wp(h->vi_out() + 1); a.out = out;
  a.B = B; a.Hi = h->dh[L]; a.Wi = h->dw[L]; a.Ci = h->dc[L];
  a.Ho = a.Hi; a.Wo = a.Wi; a.Co = h->C;
  a.w_sci = 1; a.w_sco = a.Ci; a.flip = 1;
  g_tag = "dec.out.fwd";
#ifndef KCVAE_EMU
  if (h->use_tc_out) {
    // tcgen05 implicit GEMM (tc_conv.cu): bf16 operands, fp32 accumulate in TMEM
    if (!last_is_bf16) cast_f32_to_bf16_planar(h->act_d[L], h->a_last_bf16, B, (int64_t)a.Hi * a.Wi, a.Ci, st);
    h->img_version[1] = 0;   // this image shares its buffer with the fused tail's: always rebuilt on this (unfused) path
    tc_prep_out_weights(a.w, a.Co, a.Ci, h->wimg_out, st);
    if (tc_out_conv(h->a_last_bf16, h->wimg_out, a.bias, out, B, a.Hi, a.Wi, a.Ci, a.Co, apply_sigmoid, h->tc_error, st) == 0)
      return;
    h->tc_failed = true;    // no silent downgrade: the entry point reports KCVAE_ERR_CUDA (tc_check)
    return;
  }
#endif
  conv_forward(CONV_S1, apply_sigmoid ? EPI_BIAS_SIGMOID : EPI_BIAS, a, st);
}

// encode -> reparameterize -> decode+sigmoid into h->xhat (or user buffer); z/mean/logvar in h
void run_forward(kcvae_model* h, const float* x, int B, int training, const float* eps, float* xhat, cudaStream_t st,
                 bool keep_last = true, TailOut* tail = nullptr) {
#ifndef KCVAE_EMU
  // fp32-grade (hi + lo) encoder products everywhere z / mean / logvar or a loss term leave the library; the scorer
  // (tail != nullptr: only reconstruction errors leave) runs the encoder with plain bf16 operands
  const int esplit = (tail == nullptr || keep_last || h->force_split) ? 1 : 0;
#else
  const int esplit = 0;
#endif
  run_encoder(h, x, B, st, esplit);
  const int gen = (training && !eps) ? 1 : 0;
  g_tag = "latent";
  reparameterize(h->head, B, h->latent, eps, gen, h->seed, h->rng_counter, h->z, h->mean, h->logvar,
                 gen ? h->eps_buf : nullptr, st);
  if (gen) h->rng_counter += ((uint64_t)B * h->latent + 1) / 2;
  run_decoder(h, h->z, B, 1, xhat, st, keep_last, tail);
  h->last_B = B;
}

// ------------------------------------------------------------------------- collectives
// op: 0 sum, 1 min, 2 max
int allreduce(kcvae_model* h, void* buf, int64_t count, int is_double, int op, cudaStream_t st) {
  if (h->world <= 1) return KCVAE_OK;
#ifdef KCVAE_EMU
  (void)st;
  if (!g_emu_allreduce) return fail(h, KCVAE_ERR_NCCL, "emu: no allreduce callback installed");
  g_emu_allreduce(buf, count, is_double, op);
  return KCVAE_OK;
#else
  ncclResult_t r = g_nccl.AllReduce(buf, buf, (size_t)count, is_double ? ncclFloat64 : ncclFloat32,
                                    op == 1 ? ncclMin : (op == 2 ? ncclMax : ncclSum), h->comm, st);
  if (r != ncclSuccess) return fail(h, KCVAE_ERR_NCCL, std::string("ncclAllReduce: ") + g_nccl.GetErrorString(r));
  return KCVAE_OK;
#endif
}

// ------------------------------------------------------------------------------- losses
// order `to` after everything enqueued so far on `from`
static void stream_after(kcvae_model* h, cudaStream_t from, cudaStream_t to) {
#ifndef KCVAE_EMU
  if (from == to) return;
  cudaEventRecord(h->ev_fork, from);
  cudaStreamWaitEvent(to, h->ev_fork, 0);
#else
  (void)h; (void)from; (void)to;
#endif
}

int sums_len(const kcvae_model* h) { return S_Z1 + (h->cfg.model_type == KCVAE_SINGLE ? 4 * h->latent : 4); }

// image + latent statistics (and dlogit when with_grad); all-reduced under DP
// cs: stream for the collectives (== st when nothing may overlap, e.g. kcvae_loss)
int run_stats(kcvae_model* h, const float* x, const float* xhat, int B, int tier, int with_grad, cudaStream_t st,
              cudaStream_t cs) {
  const int full = tier == KCVAE_METRICS_FULL;
  const int Bg = B * h->world;
  g_tag = "loss";
  latent_sums(h->z, h->mean, h->logvar, B, h->latent, h->cfg.model_type, h->sums, st);
  ImageStatsArgs ia{};
  ia.x = x; ia.xhat = xhat; ia.B = B; ia.P = h->P;
  ia.sums = h->sums; ia.minmax = h->minmax;
  ia.std_acc = full ? h->std_acc : nullptr;
  ia.pos_sums = (full && h->world > 1) ? h->pos_sums : nullptr;
  // tensor-core tail: d(loss)/d(logit) only as bf16 8-channel units, and the output-layer bias
  // gradient (its channel sums) straight from this pass - no fp32 copy, no colsum pass
#ifndef KCVAE_EMU
  const bool tc_tail = with_grad && ((h->use_tc_dgrad && h->use_tc_out && h->C <= 8) || h->gen_dec);
#else
  const bool tc_tail = false;
#endif
  ia.dlogit = (with_grad && !tc_tail) ? h->dlogit : nullptr;
  ia.dl8 = tc_tail ? h->dl8 : nullptr;
  ia.dbias = tc_tail ? h->gp(h->vi_out() + 1) : nullptr;
  ia.C = h->C;
  ia.grad_scale = (float)(2.0 * (double)h->lw.w_mse / ((double)Bg * (double)h->P));
  ia.want_ce = full && h->cfg.model_type == KCVAE_GLOBAL;
  ia.partial = h->dpartial;
  image_stats(ia, st);
  if (h->world > 1) {
    stream_after(h, st, cs);
    // the moment sums (<= 6 + 4 L doubles) are what latent_backward waits for: they go out alone, so the gradient path
    // never waits behind reported-only payload.  FULL tier: the per-position batch moments of x / x_hat (4 P doubles,
    // x_std_loss only) follow in their own collective; min and max share one MIN all-reduce over [min, max, -max].
    KC_TRY(allreduce(h, h->sums, (int64_t)sums_len(h), 1, 0, cs));
#ifndef KCVAE_EMU
    if (cs != st) cudaEventRecord(h->ev_sums, cs);
#endif
    if (full) {   // reported-only metrics: nothing on the gradient path waits for these
      KC_TRY(allreduce(h, h->pos_sums, 4 * h->P, 1, 0, cs));
      image_std_from_pos_sums(h->pos_sums, h->P, Bg, h->std_acc, h->dpartial, cs);
      KC_TRY(allreduce(h, h->minmax, 3, 0, 1, cs));      // [0] = global min, -[2] = global max (finalize_metrics)
    }
  }
  return KCVAE_OK;
}

void run_finalize(kcvae_model* h, int B, int tier, float* d_metrics, cudaStream_t st) {
  const int full = tier == KCVAE_METRICS_FULL;
  g_tag = "loss";
  finalize_metrics(h->sums, h->minmax, full ? h->std_acc : nullptr, B * h->world, h->latent, h->P,
                   h->cfg.model_type, h->lw, full && h->cfg.model_type == KCVAE_GLOBAL, d_metrics, st);
}

// Side stream for kernels whose results nothing on `st` waits for until the gradients are consumed (weight gradients).
// aux_fork: the side stream sees everything enqueued on st so far; aux_join: st waits for the side stream.
static cudaStream_t aux_fork(kcvae_model* h, cudaStream_t st, float** partial) {
  *partial = h->partial;
#ifndef KCVAE_EMU
  if (!h->use_aux || g_prof_on) return st;   // per-launch profiling times every kernel alone on the caller's stream
  cudaEventRecord(h->ev_aux_fork, st);
  cudaStreamWaitEvent(h->aux_stream, h->ev_aux_fork, 0);
  h->aux_dirty = true;
  *partial = h->partial2;
  return h->aux_stream;
#else
  return st;
#endif
}
// the collective stream waits for what the side stream has produced so far; the caller's stream does not
static void aux_feed(kcvae_model* h, cudaStream_t cs) {
#ifndef KCVAE_EMU
  if (!h->use_aux || !h->aux_dirty) return;
  cudaEventRecord(h->ev_aux_join, h->aux_stream);
  cudaStreamWaitEvent(cs, h->ev_aux_join, 0);
#else
  (void)h; (void)cs;
#endif
}
static void aux_join(kcvae_model* h, cudaStream_t st) {
#ifndef KCVAE_EMU
  if (!h->use_aux || !h->aux_dirty) return;
  cudaEventRecord(h->ev_aux_join, h->aux_stream);
  cudaStreamWaitEvent(st, h->ev_aux_join, 0);
  h->aux_dirty = false;
#else
  (void)h; (void)st;
#endif
}

// ------------------------------------------------------------------------------ backward
// cs != st: gradient all-reduces are issued on cs as soon as a parameter range is complete
int run_backward(kcvae_model* h, const float* x, int B, cudaStream_t st, cudaStream_t cs) {
  const int L = h->L;
  const int Bg = B * h->world;
  bool tail_s2d = false;   // d loss / d a_last lives as bf16 space-to-depth (tensor-core tail)
#ifndef KCVAE_EMU
  const bool gen_dec = h->gen_dec;
  if (gen_dec) {
    // whole decoder backward on the general engine.  Per layer: data gradient on the caller's stream (the critical chain),
    // weight + bias gradient on the side stream behind it.  Gradients travel as bf16 space-to-depth planes.
    gen_refresh(h, st);
    GenPlanes dl = pl_make(h->dl8, GEN_PLAIN, 1, 0, h->H, h->W);
    {
      const auto& g = h->gen_d[L];
      GenPlanes act = pl_act_d(h, L, 0), gout = pl_g_d(h, L);
      GenEpilogue e{};
      e.pre = GEN_PRE_NONE; e.mask = &act; e.out = &gout;
      g_tag = "dec.out.bwd";
      if (gen_conv_run(g.dgrad, dl, h->gen_wimg + g.img_dgrad, e, B, h->tc_error, "gen_dgrad", st) != 0) h->tc_failed = true;
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      g_tag = "dec.out.bwd";
      if (gen_wgrad_run(g.wgrad, act, dl, h->gp(h->vi_out()), nullptr, ax == st ? h->gen_partial : h->gen_partial2, B, h->tc_error,
                        "gen_wgrad", ax) != 0) h->tc_failed = true;      // bias gradient: image_stats
    }
    for (int l = L - 1; l >= 0; --l) {
      const auto& g = h->gen_d[l];
      const int vi = h->vi_dec_convT(l);
      GenPlanes act = pl_act_d(h, l, 0), gin = pl_g_d(h, l + 1), gout{};
      GenEpilogue e{};
      e.pre = GEN_PRE_NONE; e.mask = &act;
      if (l > 0) { gout = pl_g_d(h, l); e.out = &gout; } else e.out_f32 = h->g_act_d[0];
      g_tag = l == L - 1 ? "dec.convT_last.bwd" : (l == L - 2 ? "dec.convT.bwd" : "dec.convT_early.bwd");
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);      // the incoming gradient is complete on st: the side stream may read it
      if (gen_wgrad_run(g.wgrad, act, gin, h->gp(vi), h->gp(vi + 1), ax == st ? h->gen_partial : h->gen_partial2, B, h->tc_error,
                        "gen_wgrad", ax) != 0) h->tc_failed = true;
      g_tag = l == L - 1 ? "dec.convT_last.bwd" : (l == L - 2 ? "dec.convT.bwd" : "dec.convT_early.bwd");
      if (gen_conv_run(g.dgrad, gin, h->gen_wimg + g.img_dgrad, e, B, h->tc_error, "gen_dgrad", st) != 0) h->tc_failed = true;
    }
  }
#else
  const bool gen_dec = false;
#endif
  if (!gen_dec)
  {  // output Conv2DTranspose (s1): wgrad, bias grad, dgrad (+ReLU mask of its input)
    const int vi = h->vi_out();
    WgradArgs wa{};
    wa.P = h->dlogit; wa.Q = h->act_d[L]; wa.out = h->gp(vi); wa.partial = h->partial;
    wa.B = B; wa.Hp = h->H; wa.Wp = h->W; wa.Ca = h->C; wa.Hq = h->dh[L]; wa.Wq = h->dw[L]; wa.Cb = h->dc[L];
    wa.s = 1; wa.d = -1; wa.oy = 1; wa.ox = 1;
    wa.o_sa = wa.Cb; wa.o_sb = 1;  // [tap][out=a][in=b]
    g_tag = "dec.out.bwd";
    // The tensor-core weight gradient is enqueued AFTER the data gradient, on the side stream: the data-gradient chain is
    // the critical path; the weight-gradient kernels fill in behind it and beside the CUDA-core kernels further down.
#ifndef KCVAE_EMU
    const bool tc_w = h->use_tc_dgrad && h->use_tc_out && tc_out_wgrad_supported(h->dc[L], h->C) &&
                      h->partial_floats >= tc_out_wgrad_partial_floats(h->dc[L], h->C);
#else
    const bool tc_w = false;
#endif
    const bool tc_tail_w = h->use_tc_dgrad && h->use_tc_out && h->C <= 8;
    if (!tc_w && tc_tail_w) h->tc_failed = true;   // no fp32 dlogit exists in this mode
    if (!tc_w && !tc_tail_w) conv_wgrad(wa, st);
    const bool tc_tail = h->use_tc_dgrad && h->use_tc_out && h->C <= 8;   // bias gradient came from image_stats
    if (!tc_tail) colsum(h->dlogit, (int64_t)B * h->H * h->W, h->C, h->gp(vi + 1), h->partial, st);
    ConvArgs a{};
    a.in = h->dlogit; a.w = h->wp(vi); a.mask = h->act_d[L]; a.out = h->g_act_d[L];
    a.B = B; a.Hi = h->H; a.Wi = h->W; a.Ci = h->C; a.Ho = h->dh[L]; a.Wo = h->dw[L]; a.Co = h->dc[L];
    a.w_sci = a.Co; a.w_sco = 1; a.flip = 0;  // W[tap][co_fwd = ci'][ci_fwd = co']
    bool done = false;
#ifndef KCVAE_EMU
    if (h->use_tc_dgrad && h->use_tc_out) {  // tcgen05 paired-tap implicit GEMM (tc_conv.cu)
      if (image_stale(h, 2)) tc_prep_dgrad_weights(a.w, h->C, h->dc[L], h->wimg_dgrad, st);
      const bool s2d = h->use_tc_convT_bwd && h->use_tc_convT;
      // with the tensor-core Conv2DTranspose backward the gradient never exists in fp32: it is
      // written as bf16 space-to-depth and its channel sums (= that layer's bias gradient) come
      // out of the same epilogue
      done = tc_out_dgrad(h->dl8, h->wimg_dgrad, h->a_last_bf16, h->relu_bits_valid ? h->relu_bits : nullptr,
                          s2d ? nullptr : h->g_act_d[L], s2d ? h->g_s2d : nullptr,
                          s2d ? h->gp(h->vi_dec_convT(L - 1) + 1) : nullptr, h->partial, B, h->H, h->W, h->dc[L],
                          h->tc_error, st) == 0;
      tail_s2d = done && s2d;
    }
#endif
    if (!done && tc_tail) h->tc_failed = true;
    if (!done && !tc_tail) conv_forward(CONV_S1, EPI_MASK, a, st);
#ifndef KCVAE_EMU
    if (tc_w) {
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      g_tag = "dec.out.bwd";
      if (tc_out_wgrad(h->dl8, h->a_last_bf16, h->gp(vi), px, B, h->H, h->W, h->dc[L], h->C, h->tc_error, ax) != 0) h->tc_failed = true;
    }
#endif
  }
  for (int l = L - 1; l >= 0 && !gen_dec; --l) {  // decoder Conv2DTranspose (s2) layers
    const int vi = h->vi_dec_convT(l);
    WgradArgs wa{};
    wa.P = h->act_d[l]; wa.Q = h->g_act_d[l + 1]; wa.out = h->gp(vi); wa.partial = h->partial;
    wa.B = B; wa.Hp = h->dh[l]; wa.Wp = h->dw[l]; wa.Ca = h->dc[l];
    wa.Hq = h->dh[l + 1]; wa.Wq = h->dw[l + 1]; wa.Cb = h->dc[l + 1];
    wa.s = 2; wa.d = 1; wa.oy = 0; wa.ox = 0;
    wa.o_sa = 1; wa.o_sb = wa.Ca;  // [tap][out=b][in=a]
    g_tag = l == L - 1 ? "dec.convT_last.bwd" : (l == L - 2 ? "dec.convT.bwd" : "dec.convT_early.bwd");
#ifndef KCVAE_EMU
    if (l == L - 1 && tail_s2d) {
      if (image_stale(h, 3)) tc_prep_convT_dgrad_weights(h->wp(vi), h->dc[l + 1], h->dc[l], h->wimg_convT_dgrad, st);
      // when the layer below runs its backward on the general engine, the gradient goes straight into the planes that
      // backward reads (one 16-byte store per pixel) instead of fp32 NHWC + a repack launch
      const bool to_planes = l == 1 && h->gen_dec0 && h->pp_live && h->dc[l] <= 8 && l < (int)h->g_d_pl.size() && h->g_d_pl[l];
      GenPlanes gpl = to_planes ? pl_g_d(h, l) : GenPlanes{};
      bool ok = tc_convT_dgrad(h->g_s2d, h->wimg_convT_dgrad, h->act_d[l], to_planes ? nullptr : h->g_act_d[l], B, h->dh[l], h->dw[l],
                               h->dc[l], h->tc_error, st, to_planes ? gpl.base : nullptr, gpl.KC) == 0;
      h->g_d1_planes_only = to_planes;
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      ok = ok && tc_convT_wgrad(h->g_s2d, h->a_prev8, h->gp(vi), px, B, h->dh[l], h->dw[l], h->dc[l], h->tc_error, ax) == 0;
      if (!ok) h->tc_failed = true;
      continue;   // bias gradient was produced by tc_out_dgrad
    }
#endif
#ifndef KCVAE_EMU
    if (l == 0 && h->gen_dec0 && tail_s2d && h->pp_live) {
      // backward of the first Conv2DTranspose on the general engine: its incoming gradient (fp32 NHWC from the specialised
      // kernel behind it) is repacked as bf16 space-to-depth planes; the layer input is the hi + lo plane copy the Dense wrote
      gen_refresh(h, st);
      const auto& g = h->gen_d[0];
      GenPlanes act = pl_act_d(h, 0, 1), gin = pl_g_d(h, 1);
      if (!h->g_d1_planes_only) gen_pack_nhwc(h->g_act_d[1], B, h->dh[1], h->dw[1], h->dc[1], gin, st);
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      g_tag = "dec.convT.bwd";
      if (gen_wgrad_run(g.wgrad, act, gin, h->gp(vi), h->gp(vi + 1), ax == st ? h->gen_partial : h->gen_partial2, B, h->tc_error,
                        "gen_wgrad", ax) != 0) h->tc_failed = true;
      GenEpilogue e{};
      e.pre = GEN_PRE_NONE; e.mask = &act; e.out_f32 = h->g_act_d[0];
      g_tag = "dec.convT.bwd";
      if (gen_conv_run(g.dgrad, gin, h->gen_wimg + g.img_dgrad, e, B, h->tc_error, "gen_dgrad", st) != 0) h->tc_failed = true;
      continue;
    }
#endif
    {
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      wa.partial = px;
      conv_wgrad(wa, ax);
      colsum(h->g_act_d[l + 1], (int64_t)B * h->dh[l + 1] * h->dw[l + 1], h->dc[l + 1], h->gp(vi + 1), px, ax);
    }
    ConvArgs a{};
    a.in = h->g_act_d[l + 1]; a.w = h->wp(vi); a.mask = h->act_d[l]; a.out = h->g_act_d[l];
    a.B = B; a.Hi = h->dh[l + 1]; a.Wi = h->dw[l + 1]; a.Ci = h->dc[l + 1];
    a.Ho = h->dh[l]; a.Wo = h->dw[l]; a.Co = h->dc[l];
    a.w_sci = a.Co; a.w_sco = 1;  // W[tap][out_fwd = ci'][in_fwd = co']
    a.pad_t = 0; a.pad_l = 0;
    conv_forward(CONV_S2, EPI_MASK, a, st);
  }
  {  // decoder Dense (ReLU already folded into g_act_d[0] by the mask above)
    const int vi = h->vi_dec_dense();
    const float* G = h->g_act_d[0];
    g_tag = "dec.dense.bwd";
#ifndef KCVAE_EMU
    kcvae_model::DensePlans* dp = h->gen_dense ? dense_plans_get(h, B) : nullptr;
    if (dp) {
      dense_refresh_wT(h, st);
      // Dense backward on the engine: G -> planes [frames / 8][n][8] once; weight + bias gradient = a forward-type product over
      // them (columns = latent + a ones column, K = frames) on the side stream; data gradient = a pixel-K product of G^T and W^T
      const int K = h->latent, N = h->dec_units;
      gen_pack_rows_T(G, B, N, 0, h->gT_pl, st);
      GenPlanes gT = pl_make(h->gT_pl, GEN_PLAIN, (B + 7) / 8, 0, N / 32, 32);
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      for (size_t i = 0; i < dp->wgrad.size(); ++i) {
        unsigned char* img = dp->img + dp->off_wgrad[i];
        g_tag = "dec.dense.bwd";
        gen_conv_prep_weights(dp->wgrad[i], h->z, img, ax);
        GenEpilogue e{};
        e.pre = GEN_PRE_NONE; e.out_f32 = h->gp(vi) + (int64_t)dp->wgrad_col0[i] * N; e.dense_n = N; e.dense_ld = N; e.dense_cc = 32;
        if (gen_conv_run(dp->wgrad[i], gT, img, e, 1, h->tc_error, "gen_dense_wgrad", ax) != 0) h->tc_failed = true;
      }
      GenPlanes wT = pl_make(h->wT_pl, GEN_PLAIN, (K + 7) / 8, 1, N / 32, 32);
      g_tag = "dec.dense.bwd";
      if (gen_wgrad_run(dp->dgrad, wT, gT, h->g_z, nullptr, h->gen_partial, 1, h->tc_error, "gen_dense_dgrad", st) != 0) h->tc_failed = true;
    } else
#endif
    if (dense_wide_ok(h->z, h->wp(vi), h->gp(vi), h->gp(vi + 1), B, h->dec_units, h->latent) &&
        dense_wide_ok(G, h->g_z, h->gp(vi), nullptr, B, h->dec_units, h->latent)) {
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      dense_wide_backward(h->z, G, h->wp(vi), h->gp(vi), h->gp(vi + 1), h->g_z, h->partial, B, h->dec_units, h->latent, st, ax);
    } else {
      GemmArgs ga{};
      ga.A = h->z; ga.a_sm = 1; ga.a_sk = h->latent;
      ga.Bm = G; ga.b_sk = h->dec_units; ga.b_sn = 1;
      ga.C = h->gp(vi); ga.M = h->latent; ga.N = h->dec_units; ga.K = B; ga.partial = h->partial;
      gemm(ga, st);
      colsum(G, B, h->dec_units, h->gp(vi + 1), h->partial, st);
      GemmArgs gz{};
      gz.A = G; gz.a_sm = h->dec_units; gz.a_sk = 1;
      gz.Bm = h->wp(vi); gz.b_sk = 1; gz.b_sn = h->dec_units;
      gz.C = h->g_z; gz.M = B; gz.N = h->latent; gz.K = h->dec_units; gz.partial = h->partial;
      gemm(gz, st);
    }
  }
  if (h->world > 1) {
    // every decoder gradient is final: reduce that range (93 % of the parameters) under the encoder backward
    const int64_t off_dec = h->vars[h->vi_dec_dense()].off;
    if (cs != st) aux_feed(h, cs); else aux_join(h, st);   // weight gradients live on the side stream: only the collective waits
    stream_after(h, st, cs);
    KC_TRY(allreduce(h, h->g + off_dec, h->nparams - off_dec, 0, 0, cs));
#ifndef KCVAE_EMU
    if (cs != st) cudaStreamWaitEvent(st, h->ev_sums, 0);   // global moment sums must have landed
#endif
  }
  g_tag = "latent";
  latent_backward(h->z, h->g_z, h->sums, B, Bg, h->latent, h->cfg.model_type, h->lw, h->dhead, st);
  const float* flat_act = L > 0 ? h->act_e[L] : x;
  const float* relu_mask = L > 0 ? h->act_e[L] : nullptr;
  float* g_flat = L > 0 ? h->g_act_e[L] : nullptr;
  {  // encoder head Dense (linear)
    const int vi = h->vi_head();
    const float* hin = h->enc_dense ? h->d1 : flat_act;
    g_tag = "enc.head.bwd";
    const int kin = h->enc_dense ? h->enc_dense : h->flat;
    GemmArgs ga{};
    ga.A = hin; ga.a_sm = 1; ga.a_sk = kin;
    ga.Bm = h->dhead; ga.b_sk = 2 * h->latent; ga.b_sn = 1;
    float* px;
    cudaStream_t ax = aux_fork(h, st, &px);
    ga.C = h->gp(vi); ga.M = kin; ga.N = 2 * h->latent; ga.K = B; ga.partial = px;
    gemm(ga, ax);
    colsum(h->dhead, B, 2 * h->latent, h->gp(vi + 1), px, ax);
    float* dst = h->enc_dense ? h->g_d1 : g_flat;
    if (dst) {
      GemmArgs gi{};
      gi.A = h->dhead; gi.a_sm = 2 * h->latent; gi.a_sk = 1;
      gi.Bm = h->wp(vi); gi.b_sk = 1; gi.b_sn = 2 * h->latent;
      gi.C = dst; gi.mask = h->enc_dense ? nullptr : relu_mask;
      gi.M = B; gi.N = kin; gi.K = 2 * h->latent; gi.partial = h->partial;
      gemm(gi, st);
    }
  }
#ifndef KCVAE_EMU
  kcvae_model::EncDensePlans* edp = (h->enc_dense && h->xT_live && g_flat) ? edense_plans_get(h, B) : nullptr;
  if (edp) {
    // encoder Dense backward on the engine: dW[i][e] over the transposed activation planes the forward wrote (side stream),
    // g_flat[b][i] over the weight planes with the ReLU mask of the flattened activation; db = column sums of g (tiny)
    const int vi = h->vi_enc_dense(), F = h->ed_F, Fp = h->ed_Fp, E = h->ed_E;
    g_tag = "enc.dense.bwd";
    float* px;
    cudaStream_t ax = aux_fork(h, st, &px);
    gen_conv_prep_weights(edp->wgrad, h->g_d1, edp->img + edp->off_wgrad, ax);
    GenPlanes xT = pl_make(h->xT_pl, GEN_PLAIN, (B + 7) / 8, 1, Fp / 32, 32);
    GenEpilogue ew{};
    ew.pre = GEN_PRE_NONE; ew.out_f32 = h->gp(vi); ew.dense_n = F; ew.dense_ld = E; ew.dense_tr = 1; ew.dense_cc = 32;
    if (gen_conv_run(edp->wgrad, xT, edp->img + edp->off_wgrad, ew, 1, h->tc_error, "gen_edense_wgrad", ax) != 0) h->tc_failed = true;
    colsum(h->g_d1, B, h->enc_dense, h->gp(vi + 1), px, ax);
    g_tag = "enc.dense.bwd";
    gen_conv_prep_weights(edp->dgrad, h->g_d1, edp->img + edp->off_dgrad, st);
    GenPlanes wE = pl_make(h->wE_pl, GEN_PLAIN, (E + 7) / 8, 1, Fp / 32, 32);
    GenEpilogue eg{};
    eg.pre = GEN_PRE_NONE; eg.out_f32 = g_flat; eg.mask_f32 = relu_mask; eg.dense_n = F; eg.dense_ld = F; eg.dense_cc = 32;
    if (gen_conv_run(edp->dgrad, wE, edp->img + edp->off_dgrad, eg, 1, h->tc_error, "gen_edense_dgrad", st) != 0) h->tc_failed = true;
  } else
#endif
  if (h->enc_dense) {
    const int vi = h->vi_enc_dense();
    g_tag = "enc.dense.bwd";
    GemmArgs ga{};
    ga.A = flat_act; ga.a_sm = 1; ga.a_sk = h->flat;
    ga.Bm = h->g_d1; ga.b_sk = h->enc_dense; ga.b_sn = 1;
    float* px;
    cudaStream_t ax = aux_fork(h, st, &px);
    ga.C = h->gp(vi); ga.M = h->flat; ga.N = h->enc_dense; ga.K = B; ga.partial = px;
    gemm(ga, ax);
    colsum(h->g_d1, B, h->enc_dense, h->gp(vi + 1), px, ax);
    if (g_flat) {
      GemmArgs gi{};
      gi.A = h->g_d1; gi.a_sm = h->enc_dense; gi.a_sk = 1;
      gi.Bm = h->wp(vi); gi.b_sk = 1; gi.b_sn = h->enc_dense;
      gi.C = g_flat; gi.mask = relu_mask;
      gi.M = B; gi.N = h->flat; gi.K = h->enc_dense; gi.partial = h->partial;
      gemm(gi, st);
    }
  }
  int64_t enc_tail_floats = h->vars[h->vi_dec_dense()].off;   // encoder gradients still to be reduced at the end
  if (h->world > 1 && L > 0) {
    // the encoder Dense / head gradients (all but a few KB of the encoder range) are final here: reduce them under the
    // encoder convolutions' backward, so that only the tiny convolution range is left for the exposed collective
    const int64_t off_dense = h->vars[h->enc_dense ? h->vi_enc_dense() : h->vi_head()].off;
    if (cs != st) aux_feed(h, cs); else aux_join(h, st);
    stream_after(h, st, cs);
    KC_TRY(allreduce(h, h->g + off_dense, enc_tail_floats - off_dense, 0, 0, cs));
    enc_tail_floats = off_dense;
  }
#ifndef KCVAE_EMU
  const bool gen_enc = h->gen_enc;
  if (gen_enc && L > 0) {
    // encoder convolutions' backward on the general engine: the Dense layers hand over d loss / d act_e[L] as fp32 NHWC
    // (ReLU mask applied); from there the gradient travels as bf16 planes.  S operand of every weight gradient = the
    // space-to-depth planes the forward pass wrote (hi planes).
    gen_refresh(h, st);
    const int split = h->enc_split_live ? 1 : 0;
    g_tag = "enc.pack";
    gen_pack_nhwc(h->g_act_e[L], B, h->eh[L], h->ew[L], h->ec[L], pl_g_e(h, L), st);
    for (int l = L - 1; l >= 0; --l) {
      const auto& g = h->gen_e[l];
      const int vi = h->vi_enc_conv(l);
      GenPlanes act = l == 0 ? pl_x(h, split) : pl_act_e(h, l, split), gin = pl_g_e(h, l + 1);
      if (l == 0 && h->C == 3) act = pl_make(h->x27_pl, GEN_X27, 1, 0, h->H / 2, h->W / 2);   // the 27-value patches: one tap
      g_tag = l == 0 ? "enc.conv0.bwd" : (l == 1 ? "enc.conv1.bwd" : "enc.convN.bwd");
      if (l > 0) {
        float* px;
        cudaStream_t ax = aux_fork(h, st, &px);
        if (gen_wgrad_run(g.wgrad, act, gin, h->gp(vi), h->gp(vi + 1), ax == st ? h->gen_partial : h->gen_partial2, B, h->tc_error,
                          "gen_wgrad", ax) != 0) h->tc_failed = true;
        GenPlanes gout = pl_g_e(h, l);
        GenEpilogue e{};
        e.pre = GEN_PRE_NONE; e.mask = &act; e.out = &gout;
        g_tag = l == 1 ? "enc.conv1.bwd" : "enc.convN.bwd";
        if (gen_conv_run(g.dgrad, gin, h->gen_wimg + g.img_dgrad, e, B, h->tc_error, "gen_dgrad", st) != 0) h->tc_failed = true;
      } else {
        if (gen_wgrad_run(g.wgrad, act, gin, h->gp(vi), h->gp(vi + 1), h->gen_partial, B, h->tc_error, "gen_wgrad", st) != 0) h->tc_failed = true;
      }
    }
  }
#else
  const bool gen_enc = false;
#endif
  for (int l = L - 1; l >= 0 && !gen_enc; --l) {  // encoder Conv2D (s2) layers
    const int vi = h->vi_enc_conv(l);
    const float* in = l > 0 ? h->act_e[l] : x;
    int pt, pl;
    pad_before(h->eh[l], pt); pad_before(h->ew[l], pl);
    WgradArgs wa{};
    wa.P = h->g_act_e[l + 1]; wa.Q = in; wa.out = h->gp(vi); wa.partial = h->partial;
    wa.B = B; wa.Hp = h->eh[l + 1]; wa.Wp = h->ew[l + 1]; wa.Ca = h->ec[l + 1];
    wa.Hq = h->eh[l]; wa.Wq = h->ew[l]; wa.Cb = h->ec[l];
    wa.s = 2; wa.d = 1; wa.oy = -pt; wa.ox = -pl;
    wa.o_sa = 1; wa.o_sb = wa.Ca;  // HWIO: [tap][in=b][out=a]
    g_tag = l == 0 ? "enc.conv0.bwd" : (l == 1 ? "enc.conv1.bwd" : "enc.convN.bwd");
    wa.pcolsum = h->gp(vi + 1);   // bias gradient = column sums of P, same pass
    if (l > 0) {   // beside this layer's data gradient (and the next layer's weight gradient)
      float* px;
      cudaStream_t ax = aux_fork(h, st, &px);
      wa.partial = px;
      conv_wgrad(wa, ax);
    } else {
      conv_wgrad(wa, st);
    }
    if (l > 0) {
      ConvArgs a{};
      a.in = h->g_act_e[l + 1]; a.w = h->wp(vi); a.mask = h->act_e[l]; a.out = h->g_act_e[l];
      a.B = B; a.Hi = h->eh[l + 1]; a.Wi = h->ew[l + 1]; a.Ci = h->ec[l + 1];
      a.Ho = h->eh[l]; a.Wo = h->ew[l]; a.Co = h->ec[l];
      a.w_sci = 1; a.w_sco = a.Ci;  // W[tap][in_fwd = co'][out_fwd = ci']
      a.pad_t = pt; a.pad_l = pl;
      conv_forward(CONVT_S2, EPI_MASK, a, st);
    }
  }
  aux_join(h, st);   // every gradient is final from here on
  if (h->world > 1) {
    stream_after(h, st, cs);
    if (enc_tail_floats > 0) KC_TRY(allreduce(h, h->g, enc_tail_floats, 0, 0, cs));   // what is left of the encoder range
#ifndef KCVAE_EMU
    if (cs != st) { cudaEventRecord(h->ev_comm, cs); cudaStreamWaitEvent(st, h->ev_comm, 0); }
#endif
  }
  return KCVAE_OK;
}

int run_adam(kcvae_model* h, cudaStream_t st) {
  h->adam_t += 1;
  const double b1 = 0.9, b2 = 0.999;
  const double lr_t = (double)h->lr * std::sqrt(1.0 - std::pow(b2, (double)h->adam_t)) / (1.0 - std::pow(b1, (double)h->adam_t));
  g_tag = "optimizer";
  adam_update(h->w, h->g, h->m, h->v, h->nparams, (float)lr_t, (float)b1, (float)b2, 1e-7f, st);
  ++h->w_version;
  return KCVAE_OK;
}

int step_impl(kcvae_model* h, const float* d_x, int B, const float* d_eps, const float* d_img_noise,
              float* d_metrics, float* d_xhat, int tier, int do_update, cudaStream_t st) {
  KC_TRY(check_batch(h, B));
  KC_TRY(check_recon_shape(h));
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_bwd(h, B));
  const float* x = d_x;
  if (d_img_noise || h->train_image_noise) {  // opt-in (unreachable in the reference's train_step, SURVEY Note A)
    // caller-supplied noise (parity runs) or on-device Philox N(0, beta^2), the draw src/abstract_cvae.py:117-118 makes
    const int64_t n = (int64_t)B * h->P;
    g_tag = "noise";
    add_noise(d_x, d_img_noise, n, h->beta, h->seed, h->rng_counter, h->x_noisy, st);
    if (!d_img_noise) h->rng_counter += (uint64_t)(n + 1) / 2;
    x = h->x_noisy;
  }
#ifndef KCVAE_EMU
  // tensor-core training configuration: the four bf16 weight images of this step in one launch
  if (h->L >= 1 && h->use_tc_convT && h->use_tc_out && h->fuse_train_tail && h->use_tc_dgrad && h->use_tc_convT_bwd &&
      h->wimg_dgrad && h->wimg_convT_dgrad && tail_fusable(h, B) &&
      (h->w_external || h->img_version[0] != h->w_version || h->img_version[1] != h->w_version ||
       h->img_version[2] != h->w_version || h->img_version[3] != h->w_version)) {
    g_tag = "step";
    tc_prep_all_weights(h->wp(h->vi_dec_convT(h->L - 1)), h->wp(h->vi_out()), h->dc[h->L - 1], h->dc[h->L], h->C, h->wimg_convT,
                        h->wimg_out, h->wimg_dgrad, h->wimg_convT_dgrad, st);
    for (int k = 0; k < 4; ++k) h->img_version[k] = h->w_version;
  }
  const bool ext = h->w_external;
  if (ext) { h->gen_img_version = 0; h->wT_version = 0; h->wE_version = 0; }   // weights may have been written behind the library's back: the engine's images are rebuilt in this step
  h->w_external = false;          // the images built above are current for this step
#endif
  float* xh = d_xhat ? d_xhat : h->xhat;
  run_forward(h, x, B, 1, d_eps, xh, st);
  cudaStream_t cs = (h->world > 1 && h->comm_stream) ? h->comm_stream : st;
  KC_TRY(run_stats(h, d_x, xh, B, tier, 1, st, cs));
  KC_TRY(run_backward(h, x, B, st, cs));
  KC_TRY(tc_check(h));
#ifndef KCVAE_EMU
  h->w_external = ext;
#endif
  if (do_update) KC_TRY(run_adam(h, st));
  run_finalize(h, B, tier, d_metrics ? d_metrics : h->metrics_dev, st);
  return post(h);
}

}  // namespace

// ======================================================================================
//                                        C ABI
// ======================================================================================
extern "C" {

int kcvae_abi_version(void) { return KCVAE_ABI_VERSION; }

const char* kcvae_last_error(kcvae_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int kcvae_create(const kcvae_config* cfg, int device, kcvae_handle* out) {
  if (!cfg || !out) return fail(nullptr, KCVAE_ERR_INVALID, "null argument");
  *out = nullptr;
  if (cfg->model_type != KCVAE_GLOBAL && cfg->model_type != KCVAE_SINGLE) return fail(nullptr, KCVAE_ERR_INVALID, "unknown model_type");
  kcvae_model* h = new kcvae_model();
  h->cfg = *cfg;
  h->device = device;
  int rc = build_topology(h);
  if (rc != KCVAE_OK) { delete h; return rc; }
  h->lr = cfg->learning_rate; h->beta = cfg->beta;
  h->lw = LossWeights{cfg->kurtosis_target, cfg->w_mse, cfg->w_kurtosis, cfg->w_skew, cfg->w_z_l1_reg};
  auto bail = [&](int code) { g_create_error = h->err; kcvae_destroy(h); return code; };
  if (cudaSetDevice(device) != cudaSuccess) { h->err = "cudaSetDevice failed (no usable GPU?)"; return bail(KCVAE_ERR_CUDA); }
  if ((rc = dalloc(h, &h->w, (size_t)h->nparams)) || (rc = dalloc(h, &h->g, (size_t)h->nparams)) ||
      (rc = dalloc(h, &h->m, (size_t)h->nparams)) || (rc = dalloc(h, &h->v, (size_t)h->nparams)) ||
      (rc = dalloc(h, &h->sums, (size_t)kSumsLen)) || (rc = dalloc(h, &h->std_acc, 1)) ||
      (rc = dalloc(h, &h->minmax, 4)) || (rc = dalloc(h, &h->metrics_dev, KCVAE_NUM_METRICS)) ||
      (rc = dalloc(h, &h->dpartial, image_stats_partial_doubles())))
    return bail(rc);
  cudaMemset(h->dpartial, 0, image_stats_partial_doubles() * sizeof(double));   // holds the image_stats ticket
  cudaMemset(h->w, 0, h->nparams * sizeof(float));
  cudaMemset(h->g, 0, h->nparams * sizeof(float));
  cudaMemset(h->m, 0, h->nparams * sizeof(float));
  cudaMemset(h->v, 0, h->nparams * sizeof(float));
  cudaMemset(h->sums, 0, kSumsLen * sizeof(double));
#ifndef KCVAE_EMU
  if (cfg->precision == KCVAE_PREC_BF16_TC) {
    if ((rc = dalloc(h, &h->tc_error, 1))) return bail(rc);
    cudaMemset(h->tc_error, 0, sizeof(int));
    if (cudaMallocHost(reinterpret_cast<void**>(&h->tc_flag_host), sizeof(int)) == cudaSuccess) *h->tc_flag_host = 0;
    else h->tc_flag_host = nullptr;
  }
  if (cfg->precision == KCVAE_PREC_BF16_TC && tc_out_conv_supported(h->dc[h->L], h->C)) {
    h->use_tc_out = true;
    unsigned short* wi = nullptr;
    if ((rc = dalloc(h, &wi, tc_out_weight_image_elems(h->dc[h->L])))) return bail(rc);
    h->wimg_out = wi;
    if (h->L >= 1 && tc_convT_fwd_supported(h->dc[h->L - 1], h->dc[h->L])) {
      unsigned short* wc = nullptr;
      if ((rc = dalloc(h, &wc, tc_convT_weight_image_elems()))) return bail(rc);
      h->wimg_convT = wc;
      h->use_tc_convT = true;
      // KCVAE_TC_CONVT_FEW: 0 = fp32 CUDA-core kernel for the 32 -> few Conv2DTranspose everywhere, 1 = tensor cores for the
      // inference entry points only.  Default: inference with plain bf16 operands, training with bf16 hi + lo operand pairs
      // (fp32-grade products): with plain bf16 in front of that ReLU the decoder Dense gradient of the small golden fixture
      // moved from inside to just outside the 5e-2 relative-L2 bar the tests hold bf16 gradients to (6.1e-2).
      const char* cf = std::getenv("KCVAE_TC_CONVT_FEW");
      if (h->L >= 2 && tc_convT_few_fwd_supported(h->dc[h->L - 2], h->dc[h->L - 1]) && !(cf && cf[0] == '0')) {
        unsigned short* wf = nullptr;
        if ((rc = dalloc(h, &wf, tc_convT_few_weight_image_elems()))) return bail(rc);
        h->wimg_convT_few = wf;
        h->use_tc_convT_few = true;
        h->use_tc_convT_few_train = !(cf && cf[0] == '1');
      }
      const char* ft = std::getenv("KCVAE_FUSE_TRAIN_TAIL");   // 0 = separate convT / out-conv kernels in the training forward
      h->fuse_train_tail = !(ft && ft[0] == '0');
    }
    if (h->use_tc_convT && tc_convT_bwd_supported(h->dc[h->L - 1], h->dc[h->L], h->dh[h->L - 1], h->dw[h->L - 1]) &&
        tc_out_dgrad_supported(h->dc[h->L], h->C)) {
      unsigned short* wd2 = nullptr;
      if ((rc = dalloc(h, &wd2, tc_convT_dgrad_weight_image_elems()))) return bail(rc);
      h->wimg_convT_dgrad = wd2;
      h->use_tc_convT_bwd = true;
    }
    if (tc_out_dgrad_supported(h->dc[h->L], h->C)) {
      unsigned short* wd = nullptr;
      if ((rc = dalloc(h, &wd, tc_dgrad_weight_image_elems()))) return bail(rc);
      h->wimg_dgrad = wd;
      h->use_tc_dgrad = true;
    }
  }
  if (cfg->precision == KCVAE_PREC_BF16_TC && (rc = gen_setup(h))) return bail(rc);
#endif
  if (cfg->max_batch > 0 && (rc = ensure_fwd(h, cfg->max_batch))) return bail(rc);
  *out = h;
  return KCVAE_OK;
}

int kcvae_destroy(kcvae_handle h) {
  if (!h) return KCVAE_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
#ifndef KCVAE_EMU
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  if (h->aux_stream) { cudaStreamDestroy(h->aux_stream); cudaEventDestroy(h->ev_aux_fork); cudaEventDestroy(h->ev_aux_join); }
  if (h->partial2) cudaFree(h->partial2);
  if (h->comm_stream) { cudaStreamDestroy(h->comm_stream); cudaEventDestroy(h->ev_fork); cudaEventDestroy(h->ev_sums); cudaEventDestroy(h->ev_comm); }
#endif
#ifndef KCVAE_EMU
  if (h->copy_stream) cudaStreamDestroy(h->copy_stream);
  for (int i = 0; i < 2; ++i) { if (h->stage_free[i]) cudaEventDestroy(h->stage_free[i]); if (h->copy_done[i]) cudaEventDestroy(h->copy_done[i]); }
#endif
  for (float* p : h->act_e) if (p) cudaFree(p);
  for (float* p : h->act_d) if (p) cudaFree(p);
  for (float* p : h->g_act_e) if (p) cudaFree(p);
  for (float* p : h->g_act_d) if (p) cudaFree(p);
  float* fl[] = {h->w, h->g, h->m, h->v, h->x_stage[0], h->x_stage[1], h->d1, h->head, h->z, h->mean, h->logvar, h->eps_buf, h->xhat,
                 h->x_noisy, h->dlogit, h->g_z, h->dhead, h->g_d1, h->partial, h->err_buf, h->score_buf, h->minmax,
                 h->metrics_dev};
  for (float* p : fl) if (p) cudaFree(p);
  kc::resize_plan_free(h->resize_plan);
  for (int i = 0; i < 2; ++i) {
    if (h->u8_stage[i]) cudaFree(h->u8_stage[i]);
#ifndef KCVAE_EMU
    if (h->u8_copy_done[i]) cudaEventDestroy(h->u8_copy_done[i]);
    if (h->u8_free[i]) cudaEventDestroy(h->u8_free[i]);
#endif
  }
  if (h->a_last_bf16) cudaFree(h->a_last_bf16);
  if (h->relu_bits) cudaFree(h->relu_bits);
  if (h->wimg_out) cudaFree(h->wimg_out);
  if (h->wimg_dgrad) cudaFree(h->wimg_dgrad);
  if (h->wimg_convT) cudaFree(h->wimg_convT);
  if (h->wimg_convT_dgrad) cudaFree(h->wimg_convT_dgrad);
  if (h->g_s2d) cudaFree(h->g_s2d);
  if (h->a_prev8) cudaFree(h->a_prev8);
  if (h->a_pp_planar) cudaFree(h->a_pp_planar);
  if (h->wimg_convT_few) cudaFree(h->wimg_convT_few);
  if (h->dl8) cudaFree(h->dl8);
  if (h->tc_error) cudaFree(h->tc_error);
#ifndef KCVAE_EMU
  gen_free_plans(h->gen_e); gen_free_plans(h->gen_d);
  if (h->gen_wimg) cudaFree(h->gen_wimg);
  if (h->gen_table) cudaFree(h->gen_table);
  if (h->x_pl) cudaFree(h->x_pl);
  if (h->x27_pl) cudaFree(h->x27_pl);
  for (auto* v : {&h->act_e_pl, &h->g_e_pl, &h->act_d_pl, &h->g_d_pl}) for (void* q : *v) if (q) cudaFree(q);
  for (auto& d : h->dense_plans) dense_plans_free(d);
  for (auto& d : h->edense_plans) edense_plans_free(d);
  if (h->xT_pl) cudaFree(h->xT_pl);
  if (h->wE_pl) cudaFree(h->wE_pl);
  if (h->wT_pl) cudaFree(h->wT_pl);
  if (h->gT_pl) cudaFree(h->gT_pl);
  if (h->gen_partial) cudaFree(h->gen_partial);
  if (h->gen_partial2) cudaFree(h->gen_partial2);
#endif
  if (h->tc_flag_host) cudaFreeHost(h->tc_flag_host);
  if (h->pos_sums == h->sums + kSumsLen) h->pos_sums = nullptr;   // lives inside the sums allocation
  double* dl[] = {h->dpartial, h->sums, h->std_acc, h->pos_sums};
  for (double* p : dl) if (p) cudaFree(p);
  delete h;
  return KCVAE_OK;
}

int kcvae_num_variables(kcvae_handle h) { return h ? (int)h->vars.size() : KCVAE_ERR_INVALID; }

int64_t kcvae_param_count(kcvae_handle h) {
  if (!h) return KCVAE_ERR_INVALID;
  int64_t n = 0;
  for (const Var& v : h->vars) n += v.n;
  return n;
}

int kcvae_variable_info(kcvae_handle h, int idx, int32_t* rank, int64_t dims[4], int64_t* offset) {
  if (!h || idx < 0 || idx >= (int)h->vars.size()) return KCVAE_ERR_INVALID;
  const Var& v = h->vars[idx];
  if (rank) *rank = v.rank;
  if (dims) for (int i = 0; i < 4; ++i) dims[i] = i < v.rank ? v.dims[i] : 1;
  if (offset) *offset = v.off;
  return KCVAE_OK;
}

// h_flat is the dense concatenation of the variables (no alignment padding), n = param_count
int kcvae_set_weights(kcvae_handle h, const float* h_flat, int64_t n) {
  if (!h || !h_flat) return KCVAE_ERR_INVALID;
  if (n != kcvae_param_count(h)) return fail(h, KCVAE_ERR_INVALID, "set_weights: wrong element count");
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_CUDA(h, cudaDeviceSynchronize());
  int64_t src = 0;
  for (const Var& v : h->vars) {
    KC_CUDA(h, cudaMemcpy(h->w + v.off, h_flat + src, v.n * sizeof(float), cudaMemcpyHostToDevice));
    ++h->w_version;
    src += v.n;
  }
  return KCVAE_OK;
}

static int copy_out_flat(kcvae_handle h, const float* dev, float* h_flat, int64_t n) {
  if (!h || !h_flat) return KCVAE_ERR_INVALID;
  if (n != kcvae_param_count(h)) return fail(h, KCVAE_ERR_INVALID, "wrong element count");
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_CUDA(h, cudaDeviceSynchronize());
  int64_t dst = 0;
  for (const Var& v : h->vars) {
    KC_CUDA(h, cudaMemcpy(h_flat + dst, dev + v.off, v.n * sizeof(float), cudaMemcpyDeviceToHost));
    dst += v.n;
  }
  return KCVAE_OK;
}
int kcvae_get_weights(kcvae_handle h, float* h_flat, int64_t n) { return copy_out_flat(h, h ? h->w : nullptr, h_flat, n); }
int kcvae_get_grads(kcvae_handle h, float* h_flat, int64_t n) { return copy_out_flat(h, h ? h->g : nullptr, h_flat, n); }

float* kcvae_weights_device(kcvae_handle h) {
  if (h) h->w_external = true;   // the caller may now write weights behind the library's back
  return h ? h->w : nullptr;
}
float* kcvae_grads_device(kcvae_handle h) { return h ? h->g : nullptr; }

int kcvae_init_glorot(kcvae_handle h, uint64_t seed, void* stream) {
  if (!h) return KCVAE_ERR_INVALID;
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  ++h->w_version;
  KC_CUDA(h, cudaMemsetAsync(h->w, 0, h->nparams * sizeof(float), st));
  uint32_t sid = 0;
  for (const Var& v : h->vars) {
    ++sid;
    if (v.rank == 1) continue;  // zero biases
    double fan_in, fan_out;
    if (v.rank == 4) { fan_in = (double)v.dims[0] * v.dims[1] * v.dims[2]; fan_out = (double)v.dims[0] * v.dims[1] * v.dims[3]; }
    else { fan_in = (double)v.dims[0]; fan_out = (double)v.dims[1]; }
    glorot_fill(h->w + v.off, v.n, (float)std::sqrt(6.0 / (fan_in + fan_out)), seed, sid, st);
  }
  return post(h);
}

int kcvae_adam_reset(kcvae_handle h) {
  if (!h) return KCVAE_ERR_INVALID;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_CUDA(h, cudaDeviceSynchronize());
  KC_CUDA(h, cudaMemset(h->m, 0, h->nparams * sizeof(float)));
  KC_CUDA(h, cudaMemset(h->v, 0, h->nparams * sizeof(float)));
  h->adam_t = 0;
  return KCVAE_OK;
}

int kcvae_set_adam_state(kcvae_handle h, const float* h_m, const float* h_v, int64_t n, int64_t t) {
  if (!h || !h_m || !h_v) return KCVAE_ERR_INVALID;
  if (n != kcvae_param_count(h)) return fail(h, KCVAE_ERR_INVALID, "set_adam_state: wrong element count");
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_CUDA(h, cudaDeviceSynchronize());
  int64_t src = 0;
  for (const Var& v : h->vars) {
    KC_CUDA(h, cudaMemcpy(h->m + v.off, h_m + src, v.n * sizeof(float), cudaMemcpyHostToDevice));
    KC_CUDA(h, cudaMemcpy(h->v + v.off, h_v + src, v.n * sizeof(float), cudaMemcpyHostToDevice));
    src += v.n;
  }
  h->adam_t = t;
  return KCVAE_OK;
}

int kcvae_get_adam_state(kcvae_handle h, float* h_m, float* h_v, int64_t n, int64_t* t) {
  if (!h) return KCVAE_ERR_INVALID;
  if (h_m) KC_TRY(copy_out_flat(h, h->m, h_m, n));
  if (h_v) KC_TRY(copy_out_flat(h, h->v, h_v, n));
  if (t) *t = h->adam_t;
  return KCVAE_OK;
}

int kcvae_set_learning_rate(kcvae_handle h, float lr) { if (!h) return KCVAE_ERR_INVALID; h->lr = lr; return KCVAE_OK; }
int kcvae_set_beta(kcvae_handle h, float beta) { if (!h) return KCVAE_ERR_INVALID; h->beta = beta; return KCVAE_OK; }
int kcvae_set_train_image_noise(kcvae_handle h, int on) { if (!h) return KCVAE_ERR_INVALID; h->train_image_noise = on != 0; return KCVAE_OK; }
int kcvae_set_loss_weights(kcvae_handle h, float kurtosis_target, float w_mse, float w_kurtosis, float w_skew,
                           float w_z_l1_reg) {
  if (!h) return KCVAE_ERR_INVALID;
  h->lw = LossWeights{kurtosis_target, w_mse, w_kurtosis, w_skew, w_z_l1_reg};
  return KCVAE_OK;
}
int kcvae_seed(kcvae_handle h, uint64_t seed) {
  if (!h) return KCVAE_ERR_INVALID;
  h->seed_base = seed;
  h->seed = seed + 0x9E3779B97F4A7C15ULL * (uint64_t)h->rank;   // rank-distinct streams under data parallel
  h->rng_counter = 0;
  return KCVAE_OK;
}

// ---- data parallel ---------------------------------------------------------------------
int kcvae_comm_unique_id(void* out_id128) {
  if (!out_id128) return KCVAE_ERR_INVALID;
#ifdef KCVAE_EMU
  memset(out_id128, 0, 128);
  return KCVAE_OK;
#else
  if (!g_nccl.load(g_create_error)) return KCVAE_ERR_NCCL;
  ncclUniqueId id;
  ncclResult_t r = g_nccl.GetUniqueId(&id);
  if (r != ncclSuccess) { g_create_error = std::string("ncclGetUniqueId: ") + g_nccl.GetErrorString(r); return KCVAE_ERR_NCCL; }
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  memcpy(out_id128, &id, 128);
  return KCVAE_OK;
#endif
}

int kcvae_comm_init(kcvae_handle h, const void* id128, int rank, int world_size) {
  if (!h || !id128 || world_size < 1 || rank < 0 || rank >= world_size) return KCVAE_ERR_INVALID;
  KC_CUDA(h, cudaSetDevice(h->device));
#ifndef KCVAE_EMU
  if (!g_nccl.load(h->err)) return KCVAE_ERR_NCCL;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclResult_t r = g_nccl.CommInitRank(&h->comm, world_size, id, rank);
  if (r != ncclSuccess) return fail(h, KCVAE_ERR_NCCL, std::string("ncclCommInitRank: ") + g_nccl.GetErrorString(r));
#endif
  h->rank = rank; h->world = world_size;
  // every replica draws its own reparameterisation noise: fold the rank into the Philox key (with one key the
  // global batch would hold world-many copies of one eps tensor and bias the all-reduced kurtosis / skew moments)
  h->seed = h->seed_base + 0x9E3779B97F4A7C15ULL * (uint64_t)rank;
#ifndef KCVAE_EMU
  if (world_size > 1 && !h->comm_stream) {
    KC_CUDA(h, cudaStreamCreateWithFlags(&h->comm_stream, cudaStreamNonBlocking));
    KC_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    KC_CUDA(h, cudaEventCreateWithFlags(&h->ev_sums, cudaEventDisableTiming));
    KC_CUDA(h, cudaEventCreateWithFlags(&h->ev_comm, cudaEventDisableTiming));
  }
#endif
  if (world_size > 1) {   // per-position batch moments directly behind the moment sums: one all-reduce carries both
    KC_TRY(dalloc(h, &h->sums, (size_t)kSumsLen + (size_t)4 * h->P));
    KC_CUDA(h, cudaMemset(h->sums, 0, ((size_t)kSumsLen + (size_t)4 * h->P) * sizeof(double)));
    h->pos_sums = h->sums + kSumsLen;
  }
  return KCVAE_OK;
}

int kcvae_comm_world(kcvae_handle h) { return h ? h->world : KCVAE_ERR_INVALID; }

int kcvae_broadcast_weights(kcvae_handle h, int root, void* stream) {
  if (!h) return KCVAE_ERR_INVALID;
  if (h->world <= 1) return KCVAE_OK;
#ifdef KCVAE_EMU
  (void)root; (void)stream;
  return fail(h, KCVAE_ERR_UNSUPPORTED, "emu: broadcast not emulated");
#else
  ++h->w_version;
  ncclResult_t r = g_nccl.Broadcast(h->w, h->w, (size_t)h->nparams, ncclFloat32, root, h->comm, (cudaStream_t)stream);
  if (r != ncclSuccess) return fail(h, KCVAE_ERR_NCCL, std::string("ncclBroadcast: ") + g_nccl.GetErrorString(r));
  return KCVAE_OK;
#endif
}

// ---- forward ---------------------------------------------------------------------------
int kcvae_encode(kcvae_handle h, const float* d_x, int batch, int training, const float* d_img_noise,
                 float* d_mean, float* d_logvar, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_x || !d_mean || !d_logvar) return fail(h, KCVAE_ERR_INVALID, "encode: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  const float* x = d_x;
  if (training || d_img_noise) {  // src/abstract_cvae.py:117-118
    const int64_t n = (int64_t)batch * h->P;
    add_noise(d_x, d_img_noise, n, h->beta, h->seed, h->rng_counter, h->x_noisy, st);
    if (!d_img_noise) h->rng_counter += (uint64_t)(n + 1) / 2;
    x = h->x_noisy;
  }
  run_encoder(h, x, batch, st, 1);
  // split: z computed with eps = 0 into scratch, mean / logvar to the caller
  reparameterize(h->head, batch, h->latent, nullptr, 0, 0, 0, h->z, d_mean, d_logvar, nullptr, st);
  h->last_B = batch;
  return post(h);
}

int kcvae_reparameterize(kcvae_handle h, const float* d_mean, const float* d_logvar, int batch, int training,
                         const float* d_eps, float* d_z, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_mean || !d_logvar || !d_z) return fail(h, KCVAE_ERR_INVALID, "reparameterize: null pointer");
  KC_CUDA(h, cudaSetDevice(h->device));
  const int gen = (training && !d_eps) ? 1 : 0;
  reparam_from_parts(d_mean, d_logvar, batch, h->latent, d_eps, gen, h->seed, h->rng_counter, d_z, (cudaStream_t)stream);
  if (gen) h->rng_counter += ((uint64_t)batch * h->latent + 1) / 2;
  return post(h);
}

int kcvae_decode(kcvae_handle h, const float* d_z, int batch, int apply_sigmoid, float* d_out, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_z || !d_out) return fail(h, KCVAE_ERR_INVALID, "decode: null pointer");
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  run_decoder(h, d_z, batch, apply_sigmoid, d_out, (cudaStream_t)stream, false);
  h->last_B = batch;
  KC_TRY(tc_check(h));
  return post(h);
}

int kcvae_forward(kcvae_handle h, const float* d_x, int batch, int training, const float* d_eps, float* d_xhat,
                  float* d_z, float* d_mean, float* d_logvar, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_x || !d_xhat) return fail(h, KCVAE_ERR_INVALID, "forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  run_forward(h, d_x, batch, training, d_eps, d_xhat, st, false);
  const size_t nb = (size_t)batch * h->latent * sizeof(float);
  if (d_z) KC_CUDA(h, cudaMemcpyAsync(d_z, h->z, nb, cudaMemcpyDeviceToDevice, st));
  if (d_mean) KC_CUDA(h, cudaMemcpyAsync(d_mean, h->mean, nb, cudaMemcpyDeviceToDevice, st));
  if (d_logvar) KC_CUDA(h, cudaMemcpyAsync(d_logvar, h->logvar, nb, cudaMemcpyDeviceToDevice, st));
  KC_TRY(tc_check(h));
  return post(h);
}

// ---- loss / train ------------------------------------------------------------------------
int kcvae_loss(kcvae_handle h, const float* d_x, int batch, int training, const float* d_eps, float* d_metrics,
               float* d_xhat, int tier, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_x || !d_metrics) return fail(h, KCVAE_ERR_INVALID, "loss: null pointer");
  KC_TRY(check_recon_shape(h));
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  float* xh = d_xhat ? d_xhat : h->xhat;
#ifndef KCVAE_EMU
  h->force_split = true;
#endif
  run_forward(h, d_x, batch, training, d_eps, xh, st, false);
#ifndef KCVAE_EMU
  h->force_split = false;
#endif
  KC_TRY(tc_check(h));
  KC_TRY(run_stats(h, d_x, xh, batch, tier, 0, st, st));
  run_finalize(h, batch, tier, d_metrics, st);
  return post(h);
}

int kcvae_train_step(kcvae_handle h, const float* d_x, int batch, const float* d_eps, const float* d_img_noise,
                     float* d_metrics, float* d_xhat, int tier, void* stream) {
  if (!h || !d_x) return KCVAE_ERR_INVALID;
  return step_impl(h, d_x, batch, d_eps, d_img_noise, d_metrics, d_xhat, tier, 1, (cudaStream_t)stream);
}

int kcvae_loss_and_grads(kcvae_handle h, const float* d_x, int batch, const float* d_eps, float* d_metrics,
                         float* d_xhat, int tier, void* stream) {
  if (!h || !d_x) return KCVAE_ERR_INVALID;
  return step_impl(h, d_x, batch, d_eps, nullptr, d_metrics, d_xhat, tier, 0, (cudaStream_t)stream);
}

// ---- scoring ---------------------------------------------------------------------------
int kcvae_score(kcvae_handle h, const float* d_x, int batch, float* d_err, float* d_score, float* d_err_minmax,
                float* d_xhat, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_x || !d_score) return fail(h, KCVAE_ERR_INVALID, "score: null pointer");
  KC_TRY(check_recon_shape(h));
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  TailOut tail;
  tail.x = d_x; tail.err = d_err; tail.score = d_score; tail.err_minmax = d_err_minmax;
  // fused tail: x_hat is only written when the caller asked for it
  const bool fusable = tail_fusable(h, batch);
  float* xh = d_xhat ? d_xhat : (fusable ? nullptr : h->xhat);
  run_forward(h, d_x, batch, 0, nullptr, xh, st, false, &tail);
  KC_TRY(tc_check(h));
  if (!tail.done) {
    if (!xh) return fail(h, KCVAE_ERR_CUDA, "score: fused decoder tail unavailable (cuTensorMapEncodeTiled failed)");
    g_tag = "score";
    score(d_x, xh, batch, (int64_t)h->H * h->W, h->C, d_err, d_score, d_err_minmax, h->partial, st);
  }
  return post(h);
}

int kcvae_normalize_scores(kcvae_handle h, const float* d_err, const float* d_score, int batch, float meu,
                           float sigma, float emin, float emax, float threshold, float* d_norm, float* d_z,
                           uint8_t* d_flags, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_score) return fail(h, KCVAE_ERR_INVALID, "normalize_scores: null pointer");
  KC_CUDA(h, cudaSetDevice(h->device));
  normalize_scores(d_err, d_score, batch, (int64_t)h->H * h->W, meu, sigma, emin, emax, threshold, d_norm, d_z,
                   d_flags, (cudaStream_t)stream);
  return post(h);
}


// Brings `batch` frames from host memory into a staging buffer and returns the device pointer
// the step should read.  If the same host pointer was handed to kcvae_prefetch_host, the copy
// already runs on the copy stream and the compute stream only waits for its event.
static int acquire_host_input(kcvae_model* h, const float* h_x, int batch, cudaStream_t st, const float** d_x) {
  const size_t bytes = (size_t)batch * h->P * sizeof(float);
#ifndef KCVAE_EMU
  for (int slot = 0; slot < 2; ++slot) {
    if (h->pend_src[slot] == h_x && h->pend_batch[slot] == batch) {
      KC_CUDA(h, cudaStreamWaitEvent(st, h->copy_done[slot], 0));
      h->pend_src[slot] = nullptr;
      *d_x = h->x_in = h->x_stage[slot];
      return KCVAE_OK;
    }
  }
#endif
  const int slot = h->pend_src[0] ? 1 : 0;   // never a slot that holds an unconsumed prefetch
  if (h->pend_src[0] && h->pend_src[1]) h->pend_src[1] = nullptr;   // both claimed: drop the newer prefetch
#ifndef KCVAE_EMU
  if (h->copy_done[slot]) KC_CUDA(h, cudaStreamWaitEvent(st, h->copy_done[slot], 0));   // an abandoned prefetch may still be landing
#endif
  KC_CUDA(h, cudaMemcpyAsync(h->x_stage[slot], h_x, bytes, cudaMemcpyHostToDevice, st));
  *d_x = h->x_in = h->x_stage[slot];
  return KCVAE_OK;
}
// marks the staging slot used by the step just enqueued on `st` as reusable once the step is done
static void release_host_input(kcvae_model* h, const float* d_x, cudaStream_t st) {
#ifndef KCVAE_EMU
  const int slot = d_x == h->x_stage[1] ? 1 : 0;
  if (!h->stage_free[slot]) cudaEventCreateWithFlags(&h->stage_free[slot], cudaEventDisableTiming);
  cudaEventRecord(h->stage_free[slot], st);
#else
  (void)h; (void)d_x; (void)st;
#endif
}

// ---- host-buffer entry points -------------------------------------------------------------
int kcvae_prefetch_host(kcvae_handle h, const float* h_x, int batch) {
  KC_TRY(check_batch(h, batch));
  if (!h_x) return fail(h, KCVAE_ERR_INVALID, "prefetch_host: null pointer");
#ifdef KCVAE_EMU
  return KCVAE_OK;   // the emulator copies inline
#else
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  if (!h->copy_stream) {
    KC_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) KC_CUDA(h, cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming));
  }
  // a slot without an unconsumed prefetch; prefer the one the most recent step did not read
  int slot = h->x_in == h->x_stage[0] ? 1 : 0;
  if (h->pend_src[slot]) slot = 1 - slot;
  if (h->pend_src[slot]) return fail(h, KCVAE_ERR_INVALID, "prefetch_host: two prefetches already pending");
  if (h->stage_free[slot]) KC_CUDA(h, cudaStreamWaitEvent(h->copy_stream, h->stage_free[slot], 0));
  KC_CUDA(h, cudaMemcpyAsync(h->x_stage[slot], h_x, (size_t)batch * h->P * sizeof(float), cudaMemcpyHostToDevice, h->copy_stream));
  KC_CUDA(h, cudaEventRecord(h->copy_done[slot], h->copy_stream));
  h->pend_src[slot] = h_x; h->pend_batch[slot] = batch;
  return KCVAE_OK;
#endif
}
int kcvae_train_step_host(kcvae_handle h, const float* h_x, int batch, const float* h_eps, float* h_metrics,
                          float* h_xhat, int tier, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!h_x || !h_metrics) return fail(h, KCVAE_ERR_INVALID, "train_step_host: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_bwd(h, batch));
  const float* d_x = nullptr;
  KC_TRY(acquire_host_input(h, h_x, batch, st, &d_x));
  const float* eps = nullptr;
  if (h_eps) {
    KC_CUDA(h, cudaMemcpyAsync(h->eps_buf, h_eps, (size_t)batch * h->latent * sizeof(float), cudaMemcpyHostToDevice, st));
    eps = h->eps_buf;
  }
  KC_TRY(step_impl(h, d_x, batch, eps, nullptr, h->metrics_dev, h->xhat, tier, 1, st));
  release_host_input(h, d_x, st);
  KC_CUDA(h, cudaMemcpyAsync(h_metrics, h->metrics_dev, KCVAE_NUM_METRICS * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (h_xhat) KC_CUDA(h, cudaMemcpyAsync(h_xhat, h->xhat, (size_t)batch * h->P * sizeof(float), cudaMemcpyDeviceToHost, st));
  return tc_flag_check(h, st);
}

int kcvae_score_host(kcvae_handle h, const float* h_x, int batch, float* h_err, float* h_score, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!h_x || !h_score) return fail(h, KCVAE_ERR_INVALID, "score_host: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  const float* d_x = nullptr;
  KC_TRY(acquire_host_input(h, h_x, batch, st, &d_x));
  KC_TRY(kcvae_score(h, d_x, batch, h_err ? h->err_buf : nullptr, h->score_buf, nullptr, nullptr, stream));
  release_host_input(h, d_x, st);
  if (h_err) KC_CUDA(h, cudaMemcpyAsync(h_err, h->err_buf, (size_t)batch * h->H * h->W * sizeof(float), cudaMemcpyDeviceToHost, st));
  KC_CUDA(h, cudaMemcpyAsync(h_score, h->score_buf, (size_t)batch * sizeof(float), cudaMemcpyDeviceToHost, st));
  return tc_flag_check(h, st);
}

// ---- uint8 front end -------------------------------------------------------
static int preprocess_impl(kcvae_model* h, const uint8_t* d_frames, int batch, int in_h, int in_w, float* d_x, cudaStream_t st) {
  if (in_h <= 0 || in_w <= 0) return fail(h, KCVAE_ERR_INVALID, "preprocess_u8: invalid frame size");
  kc::ResizePlan* plan = nullptr;
  if (in_h != h->H || in_w != h->W) {
    if (!kc::resize_plan_matches(h->resize_plan, in_h, in_w, h->H, h->W, h->C)) {
      kc::resize_plan_free(h->resize_plan);
      h->resize_plan = kc::resize_plan_create(in_h, in_w, h->H, h->W, h->C);
      if (!h->resize_plan) return fail(h, KCVAE_ERR_CUDA, "preprocess_u8: could not allocate the span tables");
    }
    plan = h->resize_plan;
  }
  g_tag = "frontend";
  if (kc::preprocess_u8(d_frames, batch, plan, (int64_t)batch * h->P, d_x, st))
    return fail(h, KCVAE_ERR_CUDA, "preprocess_u8: could not allocate the row-pass scratch");
  KC_CUDA(h, cudaPeekAtLastError());
  return KCVAE_OK;
}

int kcvae_preprocess_u8(kcvae_handle h, const uint8_t* d_frames, int batch, int in_h, int in_w, float* d_x, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!d_frames || !d_x) return fail(h, KCVAE_ERR_INVALID, "preprocess_u8: null pointer");
  KC_CUDA(h, cudaSetDevice(h->device));
  return preprocess_impl(h, d_frames, batch, in_h, in_w, d_x, (cudaStream_t)stream);
}

static int u8_slot_reserve(kcvae_model* h, int slot, size_t bytes, cudaStream_t waiter) {
  if (bytes > h->u8_stage_bytes[slot]) {
    if (h->u8_stage[slot]) { KC_CUDA(h, cudaDeviceSynchronize()); cudaFree(h->u8_stage[slot]); h->u8_stage[slot] = nullptr; h->u8_stage_bytes[slot] = 0; }
    KC_CUDA(h, cudaMalloc(reinterpret_cast<void**>(&h->u8_stage[slot]), bytes));
    h->u8_stage_bytes[slot] = bytes;
  }
#ifndef KCVAE_EMU
  if (h->u8_free[slot]) KC_CUDA(h, cudaStreamWaitEvent(waiter, h->u8_free[slot], 0));   // the cast that last read this slot
#else
  (void)waiter;
#endif
  return KCVAE_OK;
}

// uint8 host frames -> device staging -> x_stage[0] (fp32 model input).  A copy started by kcvae_prefetch_host_u8
// for the same pointer is only waited for.
static int stage_host_u8(kcvae_model* h, const uint8_t* h_frames, int batch, int in_h, int in_w, cudaStream_t st, const float** d_x) {
  const size_t bytes = (size_t)batch * in_h * in_w * h->C;
  h->pend_src[0] = h->pend_src[1] = nullptr;     // a float prefetch in flight is abandoned
#ifndef KCVAE_EMU
  for (int i = 0; i < 2; ++i) if (h->copy_done[i]) KC_CUDA(h, cudaStreamWaitEvent(st, h->copy_done[i], 0));
#endif
  int slot = -1;
  for (int i = 0; i < 2; ++i)
    if (h->u8_pend_src[i] == h_frames && h->u8_pend_bytes[i] == bytes) slot = i;
  if (slot >= 0) {
#ifndef KCVAE_EMU
    KC_CUDA(h, cudaStreamWaitEvent(st, h->u8_copy_done[slot], 0));
#endif
    h->u8_pend_src[slot] = nullptr;
  } else {
    slot = h->u8_pend_src[0] ? 1 : 0;            // never a slot that holds an unconsumed prefetch
    if (h->u8_pend_src[slot]) {                  // both claimed: drop that prefetch once it has landed
#ifndef KCVAE_EMU
      KC_CUDA(h, cudaStreamWaitEvent(st, h->u8_copy_done[slot], 0));
#endif
      h->u8_pend_src[slot] = nullptr;
    }
    KC_TRY(u8_slot_reserve(h, slot, bytes, st));
    KC_CUDA(h, cudaMemcpyAsync(h->u8_stage[slot], h_frames, bytes, cudaMemcpyHostToDevice, st));
  }
  KC_TRY(preprocess_impl(h, h->u8_stage[slot], batch, in_h, in_w, h->x_stage[0], st));
#ifndef KCVAE_EMU
  if (!h->u8_free[slot]) KC_CUDA(h, cudaEventCreateWithFlags(&h->u8_free[slot], cudaEventDisableTiming));
  KC_CUDA(h, cudaEventRecord(h->u8_free[slot], st));
#endif
  h->u8_last = slot;
  *d_x = h->x_in = h->x_stage[0];
  return KCVAE_OK;
}

int kcvae_prefetch_host_u8(kcvae_handle h, const uint8_t* h_frames, int batch, int in_h, int in_w) {
  KC_TRY(check_batch(h, batch));
  if (!h_frames || in_h <= 0 || in_w <= 0) return fail(h, KCVAE_ERR_INVALID, "prefetch_host_u8: invalid arguments");
#ifdef KCVAE_EMU
  return KCVAE_OK;   // the emulator copies inline
#else
  KC_CUDA(h, cudaSetDevice(h->device));
  if (!h->copy_stream) {
    KC_CUDA(h, cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) KC_CUDA(h, cudaEventCreateWithFlags(&h->copy_done[i], cudaEventDisableTiming));
  }
  int slot = 1 - h->u8_last;
  if (h->u8_pend_src[slot]) slot = 1 - slot;
  if (h->u8_pend_src[slot]) return fail(h, KCVAE_ERR_INVALID, "prefetch_host_u8: two prefetches already pending");
  const size_t bytes = (size_t)batch * in_h * in_w * h->C;
  KC_TRY(u8_slot_reserve(h, slot, bytes, h->copy_stream));
  KC_CUDA(h, cudaMemcpyAsync(h->u8_stage[slot], h_frames, bytes, cudaMemcpyHostToDevice, h->copy_stream));
  if (!h->u8_copy_done[slot]) KC_CUDA(h, cudaEventCreateWithFlags(&h->u8_copy_done[slot], cudaEventDisableTiming));
  KC_CUDA(h, cudaEventRecord(h->u8_copy_done[slot], h->copy_stream));
  h->u8_pend_src[slot] = h_frames; h->u8_pend_bytes[slot] = bytes;
  return KCVAE_OK;
#endif
}

int kcvae_score_host_u8(kcvae_handle h, const uint8_t* h_frames, int batch, int in_h, int in_w, float* h_err, float* h_score,
                        void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!h_frames || !h_score) return fail(h, KCVAE_ERR_INVALID, "score_host_u8: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_fwd(h, batch));
  const float* d_x = nullptr;
  KC_TRY(stage_host_u8(h, h_frames, batch, in_h, in_w, st, &d_x));
  KC_TRY(kcvae_score(h, d_x, batch, h_err ? h->err_buf : nullptr, h->score_buf, nullptr, nullptr, stream));
  if (h_err) KC_CUDA(h, cudaMemcpyAsync(h_err, h->err_buf, (size_t)batch * h->H * h->W * sizeof(float), cudaMemcpyDeviceToHost, st));
  KC_CUDA(h, cudaMemcpyAsync(h_score, h->score_buf, (size_t)batch * sizeof(float), cudaMemcpyDeviceToHost, st));
  return tc_flag_check(h, st);
}

int kcvae_train_step_host_u8(kcvae_handle h, const uint8_t* h_frames, int batch, int in_h, int in_w, const float* h_eps,
                             float* h_metrics, float* h_xhat, int tier, void* stream) {
  KC_TRY(check_batch(h, batch));
  if (!h_frames || !h_metrics) return fail(h, KCVAE_ERR_INVALID, "train_step_host_u8: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  KC_CUDA(h, cudaSetDevice(h->device));
  KC_TRY(ensure_bwd(h, batch));
  const float* d_x = nullptr;
  KC_TRY(stage_host_u8(h, h_frames, batch, in_h, in_w, st, &d_x));
  const float* eps = nullptr;
  if (h_eps) {
    KC_CUDA(h, cudaMemcpyAsync(h->eps_buf, h_eps, (size_t)batch * h->latent * sizeof(float), cudaMemcpyHostToDevice, st));
    eps = h->eps_buf;
  }
  KC_TRY(step_impl(h, d_x, batch, eps, nullptr, h->metrics_dev, h->xhat, tier, 1, st));
  KC_CUDA(h, cudaMemcpyAsync(h_metrics, h->metrics_dev, KCVAE_NUM_METRICS * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (h_xhat) KC_CUDA(h, cudaMemcpyAsync(h_xhat, h->xhat, (size_t)batch * h->P * sizeof(float), cudaMemcpyDeviceToHost, st));
  return tc_flag_check(h, st);
}

// ---- introspection -------------------------------------------------------------------------
int64_t kcvae_launch_count(kcvae_handle) { return kc::g_launches; }

int64_t kcvae_debug_activation(kcvae_handle h, int which, float* h_out, int64_t capacity) {
  if (!h || h->last_B <= 0) return KCVAE_ERR_INVALID;
  const int B = h->last_B, L = h->L;
  const float* src = nullptr;
  int64_t n = 0;
  if (which >= 1 && which <= L) { src = h->act_e[which]; n = (int64_t)B * h->eh[which] * h->ew[which] * h->ec[which]; }
  else if (which >= 100 && which <= 100 + L) { int l = which - 100; src = h->act_d[l]; n = (int64_t)B * h->dh[l] * h->dw[l] * h->dc[l]; }
  else if (which == 200) { src = h->head; n = (int64_t)B * 2 * h->latent; }
  else if (which == 201) { src = h->z; n = (int64_t)B * h->latent; }
  else if (which == 300) { src = h->dlogit; n = (int64_t)B * h->P; }
  else if (which >= 301 && which <= 301 + L && (int)h->g_act_d.size() > which - 301) { int l = which - 301; src = h->g_act_d[l]; n = (int64_t)B * h->dh[l] * h->dw[l] * h->dc[l]; }
  else if (which == 350) { src = h->g_z; n = (int64_t)B * h->latent; }
  else if (which == 351) { src = h->dhead; n = (int64_t)B * 2 * h->latent; }
  else if (which >= 361 && which <= 360 + L && (int)h->g_act_e.size() > which - 360) { int l = which - 360; src = h->g_act_e[l]; n = (int64_t)B * h->eh[l] * h->ew[l] * h->ec[l]; }
#ifndef KCVAE_EMU
  // tensors that only exist as bf16 planes of the general engine are unpacked on demand
  GenPlanes pl{};
  int ph = 0, pw = 0, pc = 0;
  bool planes = false;
  if (which == 100 && h->dense_f32_skipped && h->a_pp_planar) {      // the Dense output of the last forward only exists as hi + lo planes
    pl = pl_act_d(h, 0, 1); ph = h->dh[0]; pw = h->dw[0]; pc = h->dc[0]; planes = true; src = nullptr;
  }
  if (which == 302 && h->g_d1_planes_only && (int)h->g_d_pl.size() > 1 && h->g_d_pl[1]) {   // written as planes only by the last backward
    pl = pl_g_d(h, 1); ph = h->dh[1]; pw = h->dw[1]; pc = h->dc[1]; planes = true; src = nullptr;
  }
  if (!src && n > 0 && !planes) {
    if (which >= 1 && which < L && h->gen_enc) { pl = pl_act_e(h, which, h->enc_split_live ? 1 : 0); ph = h->eh[which]; pw = h->ew[which]; pc = h->ec[which]; planes = true; }
    else if (which >= 100 && which <= 100 + L && h->gen_dec) { int l = which - 100; pl = pl_act_d(h, l, 0); ph = h->dh[l]; pw = h->dw[l]; pc = h->dc[l]; planes = true; }
    else if (which >= 302 && which <= 301 + L && h->gen_dec && (int)h->g_d_pl.size() > which - 301) { int l = which - 301; pl = pl_g_d(h, l); ph = h->dh[l]; pw = h->dw[l]; pc = h->dc[l]; planes = true; }
    else if (which >= 361 && which < 360 + L && h->gen_enc && (int)h->g_e_pl.size() > which - 360) { int l = which - 360; pl = pl_g_e(h, l); ph = h->eh[l]; pw = h->ew[l]; pc = h->ec[l]; planes = true; }
    if (planes && !pl.base) planes = false;
  }
  if (planes) {
    if (!h_out) return n;
    if (capacity < n) return fail(h, KCVAE_ERR_INVALID, "debug_activation: buffer too small");
    float* tmp = nullptr;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (cudaMalloc(reinterpret_cast<void**>(&tmp), n * sizeof(float)) != cudaSuccess) return fail(h, KCVAE_ERR_CUDA, "debug_activation: cudaMalloc failed");
    gen_unpack_nhwc(pl, B, ph, pw, pc, tmp, nullptr);
    const bool okc = cudaMemcpy(h_out, tmp, n * sizeof(float), cudaMemcpyDeviceToHost) == cudaSuccess;
    cudaFree(tmp);
    return okc ? n : fail(h, KCVAE_ERR_CUDA, "debug_activation: copy failed");
  }
#endif
  if (!src) return fail(h, KCVAE_ERR_INVALID, "debug_activation: unknown or unallocated tensor");
  if (!h_out) return n;
  if (capacity < n) return fail(h, KCVAE_ERR_INVALID, "debug_activation: buffer too small");
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  if (cudaMemcpy(h_out, src, n * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) return fail(h, KCVAE_ERR_CUDA, "debug_activation: copy failed");
  return n;
}

// 1 = tensor-core kernels active, 0 = fp32 path; negative = a tcgen05 pipeline wait expired
// (synchronises the device)
int kcvae_tc_status(kcvae_handle h) {
  if (!h) return KCVAE_ERR_INVALID;
  if (!h->tc_error) return 0;
  int flag = 0;
  cudaSetDevice(h->device);
  if (cudaDeviceSynchronize() != cudaSuccess) return fail(h, KCVAE_ERR_CUDA, "tc_status: device error");
  cudaMemcpy(&flag, h->tc_error, sizeof(int), cudaMemcpyDeviceToHost);
  if (flag) return fail(h, KCVAE_ERR_CUDA, "tcgen05 pipeline: bounded mbarrier wait expired");
#ifndef KCVAE_EMU
  return (h->use_tc_out || h->gen_enc || h->gen_dec) ? 1 : 0;
#else
  return 0;
#endif
}

// ---- per-launch timing ---------------------------------------------------------------------
int kcvae_profile_enable(int on) {
  kc::g_prof_on = on != 0;
  return KCVAE_OK;
}
// synchronises the device, then writes one line per key: "<tag>/<kernel> <launcher calls> <total ms>\n"
// and clears the records.  Returns the number of bytes needed (excluding NUL).
int64_t kcvae_profile_report(char* buf, int64_t capacity) {
#ifdef KCVAE_EMU
  if (buf && capacity > 0) buf[0] = 0;
  return 0;
#else
  cudaDeviceSynchronize();
  std::vector<std::string> keys;
  std::vector<double> ms;
  std::vector<int> cnt;
  for (auto& r : g_prof_recs) {
    float t = 0.f;
    cudaEventElapsedTime(&t, r.e0, r.e1);
    size_t i = 0;
    for (; i < keys.size(); ++i) if (keys[i] == r.key) break;
    if (i == keys.size()) { keys.push_back(r.key); ms.push_back(0); cnt.push_back(0); }
    ms[i] += t; cnt[i] += 1;
  }
  if (buf && capacity > 0) {  // a size query (buf == NULL) keeps the records
    for (auto& r : g_prof_recs) { cudaEventDestroy(r.e0); cudaEventDestroy(r.e1); }
    g_prof_recs.clear();
  }
  std::string out;
  char line[256];
  for (size_t i = 0; i < keys.size(); ++i) {
    snprintf(line, sizeof line, "%s %d %.6f\n", keys[i].c_str(), cnt[i], ms[i]);
    out += line;
  }
  if (buf && capacity > 0) {
    const size_t n = out.size() < (size_t)capacity - 1 ? out.size() : (size_t)capacity - 1;
    memcpy(buf, out.data(), n);
    buf[n] = 0;
  }
  return (int64_t)out.size();
#endif
}

// ---- layer-level hooks of the general tensor-core convolution engine (tc_gen.cu) ------------------------------------
// One product on fp32 NHWC device tensors: pack -> tcgen05 kernel -> unpack.  Parity tests drive every layer kind /
// element map / epilogue of the engine through these without building a model.
int kcvae_gen_conv_test(int kind, int w_mode, int flip, int split, int pre, int in_x3, int out_mode, int mask_mode,
                        const float* d_in, const float* d_w, const float* d_bias, const float* d_mask, float* d_out,
                        int B, int Hi, int Wi, int Ck, int Cn, void* stream) {
#ifdef KCVAE_EMU
  (void)kind; (void)w_mode; (void)flip; (void)split; (void)pre; (void)in_x3; (void)out_mode; (void)mask_mode; (void)d_in; (void)d_w;
  (void)d_bias; (void)d_mask; (void)d_out; (void)B; (void)Hi; (void)Wi; (void)Ck; (void)Cn; (void)stream;
  return fail(nullptr, KCVAE_ERR_UNSUPPORTED, "emu: tcgen05 kernels are not emulated");
#else
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_in || !d_w || !d_out || B <= 0) return fail(nullptr, KCVAE_ERR_INVALID, "gen_conv_test: invalid arguments");
  if (kind == GEN_CONV_S2 && ((Hi | Wi) & 1)) return fail(nullptr, KCVAE_ERR_INVALID, "gen_conv_test: stride-2 input must have even sizes");
  const int KCk = (Ck + 7) / 8;
  GenPlanes in{};
  in.split = split; in.KC = KCk;
  int Ho, Wo;
  if (kind == GEN_CONV_S2) { in.layout = in_x3 ? GEN_X3 : GEN_S2D; in.H = Hi / 2; in.W = Wi / 2; Ho = Hi / 2; Wo = Wi / 2; }
  else if (kind == GEN_CONVT_S2) { in.layout = GEN_PLAIN; in.H = Hi; in.W = Wi; Ho = 2 * Hi; Wo = 2 * Wi; }
  else { in.layout = GEN_PLAIN; in.H = Hi; in.W = Wi; Ho = Hi; Wo = Wi; }
  GenConvSpec sp{};
  sp.kind = kind; sp.in_layout = in.layout; sp.Ck = Ck; sp.Cn = Cn; sp.KCk = KCk; sp.w_mode = w_mode; sp.flip = flip; sp.split = split;
  sp.Hg = kind == GEN_CONVT_S2 ? Hi : Ho; sp.Wg = kind == GEN_CONVT_S2 ? Wi : Wo;
  const char* why = "";
  GenConvPlan* plan = gen_conv_plan_create(sp, &why);
  if (!plan) return fail(nullptr, KCVAE_ERR_UNSUPPORTED, std::string("gen_conv_test: ") + why);
  void *d_planes = nullptr, *d_img = nullptr, *d_oplanes = nullptr, *d_mplanes = nullptr;
  int* d_err = nullptr;
  int rc = KCVAE_OK;
  const int KCo = (gen_conv_Cop(plan)) / 8;
  GenPlanes out{}, mask{};
  out.layout = out_mode == 2 ? GEN_S2D : GEN_PLAIN; out.KC = KCo; out.split = 1;
  out.H = out_mode == 2 ? Ho / 2 : Ho; out.W = out_mode == 2 ? Wo / 2 : Wo;
  mask = out; mask.split = 0; mask.layout = mask_mode == 2 ? GEN_S2D : GEN_PLAIN;
  mask.H = mask_mode == 2 ? Ho / 2 : Ho; mask.W = mask_mode == 2 ? Wo / 2 : Wo;
  auto done = [&](int code, const char* msg) {
    cudaStreamSynchronize(st);
    if (d_planes) cudaFree(d_planes); if (d_img) cudaFree(d_img); if (d_oplanes) cudaFree(d_oplanes);
    if (d_mplanes) cudaFree(d_mplanes); if (d_err) cudaFree(d_err);
    gen_conv_plan_free(plan);
    return code == KCVAE_OK ? KCVAE_OK : fail(nullptr, code, msg);
  };
  if (cudaMalloc(&d_planes, in.units(B) * 16) != cudaSuccess || cudaMalloc(&d_img, gen_conv_weight_image_bytes(plan) + 16) != cudaSuccess ||
      cudaMalloc(&d_oplanes, out.units(B) * 16) != cudaSuccess || cudaMalloc(&d_mplanes, mask.units(B) * 16) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&d_err), sizeof(int)) != cudaSuccess)
    return done(KCVAE_ERR_CUDA, "gen_conv_test: cudaMalloc failed");
  cudaMemsetAsync(d_err, 0, sizeof(int), st);
  cudaMemsetAsync(d_oplanes, 0, out.units(B) * 16, st);
  in.base = d_planes; out.base = d_oplanes; mask.base = d_mplanes;
  if (in.layout == GEN_X3) gen_pack_x3(d_in, B, Hi, Wi, split, d_planes, st);
  else if (in.layout == GEN_S2D) { GenPlanes full = in; gen_pack_nhwc(d_in, B, Hi, Wi, Ck, full, st); }
  else gen_pack_nhwc(d_in, B, Hi, Wi, Ck, in, st);
  if (d_mask && mask_mode) gen_pack_nhwc(d_mask, B, Ho, Wo, Cn, mask, st);
  gen_conv_prep_weights(plan, d_w, d_img, st);
  GenEpilogue e{};
  e.pre = pre; e.bias = d_bias;
  e.mask = (d_mask && mask_mode) ? &mask : nullptr;
  e.mask_f32 = (d_mask && !mask_mode) ? d_mask : nullptr;
  e.out = out_mode ? &out : nullptr;
  e.out_f32 = out_mode ? nullptr : d_out;
  rc = gen_conv_run(plan, in, d_img, e, B, d_err, "gen_conv_test", st);
  if (rc != 0) return done(KCVAE_ERR_CUDA, "gen_conv_test: launcher failed (tensor map / plane count)");
  if (out_mode) gen_unpack_nhwc(out, B, Ho, Wo, Cn, d_out, st);
  int flag = 0;
  cudaMemcpyAsync(&flag, d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_conv_test: kernel failed");
  if (flag) return done(KCVAE_ERR_CUDA, "gen_conv_test: bounded mbarrier wait expired");
  return done(KCVAE_OK, "");
#endif
}

int kcvae_gen_wgrad_test(int kind, int w_mode, int flip, int s_x3, const float* d_s, const float* d_u, float* d_dW, float* d_db,
                         int B, int Hs, int Ws, int Cs, int Cu, void* stream) {
#ifdef KCVAE_EMU
  (void)kind; (void)w_mode; (void)flip; (void)s_x3; (void)d_s; (void)d_u; (void)d_dW; (void)d_db; (void)B; (void)Hs; (void)Ws; (void)Cs;
  (void)Cu; (void)stream;
  return fail(nullptr, KCVAE_ERR_UNSUPPORTED, "emu: tcgen05 kernels are not emulated");
#else
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_s || !d_u || !d_dW || B <= 0) return fail(nullptr, KCVAE_ERR_INVALID, "gen_wgrad_test: invalid arguments");
  GenPlanes S{}, U{};
  S.KC = (Cs + 7) / 8; U.KC = (Cu + 7) / 8;
  int Hu, Wu;          // full-resolution dims of the gradient tensor as fp32 NHWC
  if (kind == GEN_CONV_S2) {          // S: layer input at (Hs, Ws), stored S2D / X3; U: gradient at (Hs/2, Ws/2), PLAIN
    if ((Hs | Ws) & 1) return fail(nullptr, KCVAE_ERR_INVALID, "gen_wgrad_test: stride-2 input must have even sizes");
    S.layout = s_x3 == 2 ? GEN_X27 : (s_x3 ? GEN_X3 : GEN_S2D); S.H = Hs / 2; S.W = Ws / 2;
    U.layout = GEN_PLAIN; U.H = Hs / 2; U.W = Ws / 2; Hu = Hs / 2; Wu = Ws / 2;
  } else if (kind == GEN_CONVT_S2) {  // S: layer input at (Hs, Ws), PLAIN; U: gradient at (2Hs, 2Ws), stored S2D
    S.layout = GEN_PLAIN; S.H = Hs; S.W = Ws;
    U.layout = GEN_S2D; U.H = Hs; U.W = Ws; Hu = 2 * Hs; Wu = 2 * Ws;
  } else {
    S.layout = GEN_PLAIN; S.H = Hs; S.W = Ws; U.layout = GEN_PLAIN; U.H = Hs; U.W = Ws; Hu = Hs; Wu = Ws;
  }
  GenWgradSpec sp{};
  sp.kind = kind; sp.flip = flip; sp.s_layout = S.layout; sp.s_KC = S.KC; sp.u_layout = U.layout; sp.u_KC = U.KC;
  sp.Cs = Cs; sp.Cu = Cu; sp.w_mode = w_mode; sp.Hg = U.H; sp.Wg = U.W;
  const char* why = "";
  GenWgradPlan* plan = gen_wgrad_plan_create(sp, &why);
  if (!plan) return fail(nullptr, KCVAE_ERR_UNSUPPORTED, std::string("gen_wgrad_test: ") + why);
  void *ds = nullptr, *du = nullptr;
  float* part = nullptr;
  int* d_err = nullptr;
  auto done = [&](int code, const char* msg) {
    cudaStreamSynchronize(st);
    if (ds) cudaFree(ds); if (du) cudaFree(du); if (part) cudaFree(part); if (d_err) cudaFree(d_err);
    gen_wgrad_plan_free(plan);
    return code == KCVAE_OK ? KCVAE_OK : fail(nullptr, code, msg);
  };
  if (cudaMalloc(&ds, S.units(B) * 16) != cudaSuccess || cudaMalloc(&du, U.units(B) * 16) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&part), gen_wgrad_partial_floats(plan) * sizeof(float)) != cudaSuccess ||
      cudaMalloc(reinterpret_cast<void**>(&d_err), sizeof(int)) != cudaSuccess)
    return done(KCVAE_ERR_CUDA, "gen_wgrad_test: cudaMalloc failed");
  cudaMemsetAsync(d_err, 0, sizeof(int), st);
  S.base = ds; U.base = du;
  if (s_x3 == 2) {            // X27 patch planes (written beside the X3 planes by the same packer)
    void* tmp = nullptr;
    if (cudaMalloc(&tmp, (size_t)B * 2 * (Hs / 2) * (Ws / 2) * 16) != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_wgrad_test: cudaMalloc failed");
    gen_pack_x3(d_s, B, Hs, Ws, 0, tmp, st, ds);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
  } else if (S.layout == GEN_X3) gen_pack_x3(d_s, B, Hs, Ws, 0, ds, st);
  else gen_pack_nhwc(d_s, B, Hs, Ws, Cs, S, st);
  gen_pack_nhwc(d_u, B, Hu, Wu, Cu, U, st);
  if (gen_wgrad_run(plan, S, U, d_dW, d_db, part, B, d_err, "gen_wgrad_test", st) != 0)
    return done(KCVAE_ERR_CUDA, "gen_wgrad_test: launcher failed (tensor map / plane count)");
  int flag = 0;
  cudaMemcpyAsync(&flag, d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_wgrad_test: kernel failed");
  if (flag) return done(KCVAE_ERR_CUDA, "gen_wgrad_test: bounded mbarrier wait expired");
  return done(KCVAE_OK, "");
#endif
}

// Dense-layer products of the engine on fp32 device matrices (tests): A is [R][N] row-major (N % 32 == 0).
//   mode 0  out[c][n] = act(bias[n] + sum_r A[r][n] * Bm[c][r])     Dense forward      (A = W [K][N], Bm = z [batch][K])
//   mode 1  out[c][n] = sum_r A[r][n] * Bm[r][c], out[C][n] = sum_r A[r][n]   weight + bias gradient (A = G [batch][N], Bm = z [batch][K])
//   mode 2  out[u][s] = sum_n A[u][n] * Bm[s][n]                    data gradient      (A = G [batch][N], Bm = W [K][N]), R = rows of A, C = rows of Bm
int kcvae_gen_dense_test(int mode, int split, int relu, const float* d_a, const float* d_b, const float* d_bias, float* d_out,
                         int R, int N, int C, void* stream) {
#ifdef KCVAE_EMU
  (void)mode; (void)split; (void)relu; (void)d_a; (void)d_b; (void)d_bias; (void)d_out; (void)R; (void)N; (void)C; (void)stream;
  return fail(nullptr, KCVAE_ERR_UNSUPPORTED, "emu: tcgen05 kernels are not emulated");
#else
  cudaStream_t st = (cudaStream_t)stream;
  if (!d_a || !d_b || !d_out || R <= 0 || C <= 0 || N <= 0 || N % 32) return fail(nullptr, KCVAE_ERR_INVALID, "gen_dense_test: invalid arguments");
  const char* why = "";
  void *pa = nullptr, *pb = nullptr, *img = nullptr;
  float *src = nullptr, *part = nullptr;
  int32_t* tab = nullptr;
  int* d_err = nullptr;
  GenConvPlan* cp = nullptr;
  GenWgradPlan* wp = nullptr;
  auto done = [&](int code, const char* msg) {
    cudaStreamSynchronize(st);
    for (void* q : {pa, pb, img, (void*)src, (void*)part, (void*)tab, (void*)d_err}) if (q) cudaFree(q);
    gen_conv_plan_free(cp); gen_wgrad_plan_free(wp);
    return code == KCVAE_OK ? KCVAE_OK : fail(nullptr, code, msg);
  };
  if (cudaMalloc(reinterpret_cast<void**>(&d_err), sizeof(int)) != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_dense_test: cudaMalloc failed");
  cudaMemsetAsync(d_err, 0, sizeof(int), st);
  const int KCa = (R + 7) / 8;
  GenPlanes A = pl_make(nullptr, GEN_PLAIN, KCa, mode == 0 ? split : 0, N / 32, 32);
  if (cudaMalloc(&pa, A.units(1) * 16) != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_dense_test: cudaMalloc failed");
  A.base = pa;
  gen_pack_rows_T(d_a, R, N, A.split, pa, st);
  int rc = 0;
  if (mode == 0 || mode == 1) {
    GenConvSpec sp{};
    sp.kind = GEN_DENSE; sp.in_layout = GEN_PLAIN; sp.Ck = R; sp.KCk = KCa; sp.Hg = N / 32; sp.Wg = 32; sp.split = mode == 0 ? split : 0;
    if (mode == 0) { sp.Cn = C; sp.w_mode = 1; }
    else { sp.Cn = C + 1; sp.w_mode = 0; sp.w_stride = C; sp.ones_col1 = C + 1; sp.ones_src = R * C; }
    cp = gen_conv_plan_create(sp, &why);
    if (!cp) return done(KCVAE_ERR_UNSUPPORTED, why);
    // "weight" source: Bm followed by a 1.0f
    const size_t nb = (size_t)(mode == 0 ? C * R : R * C);
    const float one = 1.0f;
    if (cudaMalloc(reinterpret_cast<void**>(&src), (nb + 4) * sizeof(float)) != cudaSuccess ||
        cudaMalloc(&img, gen_conv_weight_image_bytes(cp) + 16) != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_dense_test: cudaMalloc failed");
    cudaMemcpyAsync(src, d_b, nb * sizeof(float), cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(src + nb, &one, sizeof(float), cudaMemcpyHostToDevice, st);
    cudaStreamSynchronize(st);
    gen_conv_prep_weights(cp, src, img, st);
    GenEpilogue e{};
    e.pre = mode == 0 ? (relu ? GEN_PRE_BIAS_RELU : GEN_PRE_BIAS) : GEN_PRE_NONE;
    e.bias = mode == 0 ? d_bias : nullptr;
    e.out_f32 = d_out; e.dense_n = N; e.dense_ld = N; e.dense_cc = 32;
    rc = gen_conv_run(cp, A, img, e, 1, d_err, "gen_dense_test", st);
  } else {
    const int KCb = (C + 7) / 8;
    GenPlanes Bp = pl_make(nullptr, GEN_PLAIN, KCb, 0, N / 32, 32);
    if (cudaMalloc(&pb, Bp.units(1) * 16) != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_dense_test: cudaMalloc failed");
    Bp.base = pb;
    gen_pack_rows_T(d_b, C, N, 0, pb, st);
    GenWgradSpec sp{};
    sp.kind = GEN_DENSE; sp.s_layout = GEN_PLAIN; sp.s_KC = KCb; sp.u_layout = GEN_PLAIN; sp.u_KC = KCa; sp.Cs = C; sp.Cu = R; sp.w_mode = 1;
    sp.Hg = N / 32; sp.Wg = 32;
    wp = gen_wgrad_plan_create(sp, &why);
    if (!wp) return done(KCVAE_ERR_UNSUPPORTED, why);
    if (cudaMalloc(reinterpret_cast<void**>(&part), gen_wgrad_partial_floats(wp) * sizeof(float)) != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_dense_test: cudaMalloc failed");
    rc = gen_wgrad_run(wp, Bp, A, d_out, nullptr, part, 1, d_err, "gen_dense_test", st);
  }
  if (rc != 0) return done(KCVAE_ERR_CUDA, "gen_dense_test: launcher failed");
  int flag = 0;
  cudaMemcpyAsync(&flag, d_err, sizeof(int), cudaMemcpyDeviceToHost, st);
  if (cudaStreamSynchronize(st) != cudaSuccess || cudaGetLastError() != cudaSuccess) return done(KCVAE_ERR_CUDA, "gen_dense_test: kernel failed");
  if (flag) return done(KCVAE_ERR_CUDA, "gen_dense_test: bounded mbarrier wait expired");
  return done(KCVAE_OK, "");
#endif
}

// Host-side plan of one product of the general engine as a flat int32 array (no GPU needed): the CPU test-suite interprets
// the MMA list / gather table / scatter table with numpy (tests/engine_sim.py) and checks the planner against the oracle.
// which 0: spec = {kind, in_layout, Ck, Cn, KCk, w_mode, flip, split, w_stride, ones_col1, ones_src, w_col0, Hg, Wg}
// which 1: spec = {kind, flip, s_layout, s_KC, u_layout, u_KC, Cs, Cu, w_mode, Hg, Wg}
int64_t kcvae_gen_plan_dump(int which, const int32_t* spec, int nspec, int32_t* out, int64_t capacity) {
#ifdef KCVAE_EMU
  (void)which; (void)spec; (void)nspec; (void)out; (void)capacity;
  return fail(nullptr, KCVAE_ERR_UNSUPPORTED, "emu: the engine's planners live in the CUDA library");
#else
  const char* why = "";
  if (!spec) return fail(nullptr, KCVAE_ERR_INVALID, "gen_plan_dump: null spec");
  if (which == 0) {
    if (nspec < 14) return fail(nullptr, KCVAE_ERR_INVALID, "gen_plan_dump: 14 spec entries expected");
    GenConvSpec s{};
    s.kind = spec[0]; s.in_layout = spec[1]; s.Ck = spec[2]; s.Cn = spec[3]; s.KCk = spec[4]; s.w_mode = spec[5]; s.flip = spec[6];
    s.split = spec[7]; s.w_stride = spec[8]; s.ones_col1 = spec[9]; s.ones_src = spec[10]; s.w_col0 = spec[11]; s.Hg = spec[12]; s.Wg = spec[13];
    GenConvPlan* p = gen_conv_plan_create(s, &why, false);
    if (!p) return fail(nullptr, KCVAE_ERR_UNSUPPORTED, std::string("gen_plan_dump: ") + why);
    const int64_t n = gen_conv_plan_dump(p, out, capacity);
    gen_conv_plan_free(p);
    return n;
  }
  if (nspec < 11) return fail(nullptr, KCVAE_ERR_INVALID, "gen_plan_dump: 11 spec entries expected");
  GenWgradSpec s{};
  s.kind = spec[0]; s.flip = spec[1]; s.s_layout = spec[2]; s.s_KC = spec[3]; s.u_layout = spec[4]; s.u_KC = spec[5]; s.Cs = spec[6];
  s.Cu = spec[7]; s.w_mode = spec[8]; s.Hg = spec[9]; s.Wg = spec[10];
  if (nspec > 11) s.split_dense = spec[11];
  GenWgradPlan* p = gen_wgrad_plan_create(s, &why, false);
  if (!p) return fail(nullptr, KCVAE_ERR_UNSUPPORTED, std::string("gen_plan_dump: ") + why);
  const int64_t n = gen_wgrad_plan_dump(p, out, capacity);
  gen_wgrad_plan_free(p);
  return n;
#endif
}

#ifdef KCVAE_EMU
// tests only: route the data-parallel all-reduce through a host callback (gloo in pytest)
void kcvae_emu_set_allreduce(void (*fn)(void*, int64_t, int, int)) { g_emu_allreduce = fn; }
#endif

}  // extern "C"
