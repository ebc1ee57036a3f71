// extern "C" surface of libpdr.so (declared in include/pdr.h).
#include "pdr.h"
#include "common.cuh"
#include "conv_tc.h"

using namespace pdr;

extern "C" {

int pdr_version(void) { return PDR_VERSION; }
const char* pdr_last_error(void) { return get_error(); }
unsigned long long pdr_launch_count(void) { return g_launch_count; }

int pdr_conv_tc(const void* x1, const void* x2, const void* w, const float* bias,
                const void* residual, void* out, int B, int H, int W, int C1, int C2, int Cout,
                int taps, int bn, void* stream) {
  PDR_CHECK_ARG(x1 && w && out, "pdr_conv_tc: null pointer");
  PDR_CHECK_ARG(B > 0 && H > 0 && W > 0, "pdr_conv_tc: empty shape");
  PDR_CHECK_ARG(Cout % 128 == 0, "pdr_conv_tc: Cout (%d) must be a multiple of 128", Cout);
  if (bn == 0) bn = conv_tc_pick_bn(B, H, W, Cout);
  ConvTensorMap ma1, ma2, mw;
  PDR_TRY(conv_tc_make_act_map(&ma1, x1, B, H, W, C1));
  if (C2 > 0) {
    PDR_CHECK_ARG(x2 != nullptr, "pdr_conv_tc: x2 is null but C2=%d", C2);
    PDR_TRY(conv_tc_make_act_map(&ma2, x2, B, H, W, C2));
  }
  PDR_TRY(conv_tc_make_weight_map(&mw, w, Cout, taps * (C1 + C2), bn));
  return conv_tc_launch(&ma1, C2 > 0 ? &ma2 : nullptr, &mw, bn, B, H, W, C1, C2, Cout, taps, bias,
                        (const __half*)residual, (__half*)out, (cudaStream_t)stream);
}

}  // extern "C"
