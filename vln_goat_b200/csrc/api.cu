// extern "C" entry points of libgoat_sm100 (see include/goat_sm100.h).
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace goat {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("GOAT_PDL");
    v = (e && *e == '0') ? 0 : 1;
  }
  return v != 0;
}

bool gemm_umma_eligible(const goat_gemm_args& a);
int gemm_umma(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream);
int gemm_simt(const goat_gemm_args& a, const EpiParams& ep, cudaStream_t stream);

}  // namespace goat

using namespace goat;

extern "C" {

int goat_version(void) { return 100; }

const char* goat_last_error(void) { return g_err; }

int goat_device_supported(void) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 0;
  cudaDeviceProp p;
  if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
  return p.major == 10 ? 1 : 0;
}

int goat_gemm(const goat_gemm_args* a, goat_stream_t stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  GOAT_CHECK(a != nullptr, "goat_gemm: null args");
  GOAT_CHECK(a->M >= 0 && a->N >= 0 && a->K >= 0, "goat_gemm: negative dimension");
  if (a->M == 0 || a->N == 0) return GOAT_OK;
  GOAT_CHECK(a->K > 0, "goat_gemm: K must be > 0");
  GOAT_CHECK(a->A && a->B && a->out, "goat_gemm: null A/B/out");
  GOAT_CHECK(a->dtype == GOAT_F32 || a->dtype == GOAT_F16 || a->dtype == GOAT_BF16, "goat_gemm: bad dtype %d", a->dtype);
  GOAT_CHECK(a->out_dtype == GOAT_F32 || a->out_dtype == a->dtype, "goat_gemm: out_dtype must be F32 or the operand dtype");
  GOAT_CHECK(a->act >= GOAT_ACT_NONE && a->act <= GOAT_ACT_TANH, "goat_gemm: bad act %d", a->act);
  GOAT_CHECK(!((a->act == GOAT_ACT_DGELU || a->act == GOAT_ACT_DRELU) && !a->aux_in), "goat_gemm: DGELU/DRELU need aux_in");
  GOAT_CHECK(a->drop_p >= 0.0f && a->drop_p < 1.0f, "goat_gemm: drop_p out of range");
  GOAT_CHECK(a->lda >= (a->a_mn_major ? a->M : a->K), "goat_gemm: lda too small");
  GOAT_CHECK(a->ldb >= (a->b_mn_major ? a->N : a->K), "goat_gemm: ldb too small");
  GOAT_CHECK(a->ldc >= a->N, "goat_gemm: ldc too small");
  GOAT_CHECK(!(a->out2 && a->dtype == GOAT_F32 && a->out_dtype == GOAT_F32 && a->out2 == a->out), "goat_gemm: out2 aliases out");
  EpiParams ep;
  ep.bias = a->bias;
  ep.res = a->res;
  ep.aux_in = a->aux_in;
  ep.aux_out = a->aux_out;
  ep.out = a->out;
  ep.out2 = a->out2;
  ep.ldc = a->ldc;
  ep.ldres = a->ldres;
  ep.ldaux = a->ldaux;
  ep.ldc2 = a->ldc2;
  ep.act = a->act;
  ep.out_f32 = (a->out_dtype == GOAT_F32) ? 1 : 0;
  ep.accumulate = a->accumulate ? 1 : 0;
  GOAT_CHECK(!a->accumulate || (a->out_dtype == GOAT_F32 && !a->bias && !a->res && !a->out2 && a->act == GOAT_ACT_NONE &&
                                a->drop_p == 0.0f && !a->aux_out),
             "goat_gemm: accumulate mode takes an fp32 output and no bias/res/act/dropout/out2");
  ep.alpha = a->alpha;
  ep.drop_p = a->drop_p;
  ep.drop_seed = a->drop_seed;
  ep.drop_seed_ptr = reinterpret_cast<const unsigned long long*>(a->drop_seed_ptr);
  ep.drop_row0 = 0;
  GOAT_CHECK(!a->B_lo || (a->dtype != GOAT_F32 && !a->b_mn_major && !a->accumulate && aligned16(a->B_lo)),
             "goat_gemm: B_lo needs 16-bit operands, a K-major B, no accumulate mode and a 16-byte aligned pointer");
  if (!a->force_simt && gemm_umma_eligible(*a)) {
    GOAT_CHECK(aligned16(a->out) && (!a->out2 || aligned16(a->out2)) && (!a->res || aligned16(a->res)) &&
                   (!a->aux_in || aligned16(a->aux_in)) && (!a->aux_out || aligned16(a->aux_out)) &&
                   (!a->bias || aligned16(a->bias)),
               "goat_gemm: tensor base pointers must be 16-byte aligned");
    return gemm_umma(*a, ep, stream);
  }
  GOAT_CHECK(!a->B_lo, "goat_gemm: B_lo is a tcgen05-path operand (K >= 16, leading dimensions multiples of 8)");
  return gemm_simt(*a, ep, stream);
}

}  // extern "C"
