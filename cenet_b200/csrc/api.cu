// Library-level plumbing: error string, launch counter, GEMM dispatch.
#include "common.cuh"
#include <cstdarg>
#include <cstdio>

static thread_local char g_err[512] = "";
std::atomic<long long> g_cenet_launches{0};

void cenet_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* cenet_last_error(void) { return g_err; }
extern "C" int cenet_abi_version(void) { return 1; }
extern "C" long long cenet_launch_count(void) { return g_cenet_launches.load(); }

extern "C" int cenet_gemm(const cenet_gemm_args* a, cenet_stream_t s) {
  CENET_REQUIRE(a != nullptr, "cenet_gemm: null args");
  CENET_REQUIRE(a->M >= 0 && a->N > 0 && a->K > 0, "cenet_gemm: bad shape M=%d N=%d K=%d", a->M, a->N, a->K);
  if (a->M == 0) return 0;
  CENET_REQUIRE(a->A && a->Wt && a->C, "cenet_gemm: null operand");
  CENET_REQUIRE(a->batch >= 1 && a->batch_inner >= 1, "cenet_gemm: batch must be >= 1");
  if (a->conv) {
    CENET_REQUIRE(a->K == a->KH * a->KW * a->Cin, "cenet_gemm(conv): K=%d != KH*KW*Cin=%d", a->K, a->KH * a->KW * a->Cin);
    CENET_REQUIRE(a->M == a->Bimg * a->Ho * a->Wo, "cenet_gemm(conv): M=%d != B*Ho*Wo", a->M);
    CENET_REQUIRE(a->batch == 1, "cenet_gemm(conv): batch must be 1");
  }
  int impl = a->impl;
  if (impl == CENET_GEMM_AUTO)
    impl = cenet_gemm_tc_eligible(a) ? CENET_GEMM_TCGEN05 : (cenet_gemm_mma_eligible(a) ? CENET_GEMM_MMA : CENET_GEMM_SIMT);
  if (impl == CENET_GEMM_MMA) {
    CENET_REQUIRE(cenet_gemm_mma_eligible(a), "cenet_gemm: the mma.sync path needs bf16 A and W, no conv, no k_scale");
    return cenet_gemm_mma(a, to_stream(s));
  }
  if (impl == CENET_GEMM_TCGEN05) {
    CENET_REQUIRE(cenet_gemm_tc_eligible(a), "cenet_gemm: tcgen05 path needs bf16 K-major operands, K%%8==0, "
                  "16-byte aligned rows (M=%d N=%d K=%d conv=%d)", a->M, a->N, a->K, a->conv);
    return cenet_gemm_tc(a, to_stream(s));
  }
  return cenet_gemm_simt(a, to_stream(s));
}
