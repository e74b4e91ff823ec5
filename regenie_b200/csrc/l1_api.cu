// Level-1 ridge, LOCO assembly and Step-2 entry points of the C ABI (include/rg_b200.h).
#include "context.cuh"

using namespace rg;

#define RG_API_BEGIN try {
#define RG_API_END                         \
  }                                        \
  catch (const rg::Error& e) {             \
    rg::set_last_error(e.msg);             \
    return 1;                              \
  }                                        \
  catch (const std::exception& e) {        \
    rg::set_last_error(e.what());          \
    return 1;                              \
  }                                        \
  return 0;

extern "C" {

int rg_l1_fit(rg_handle h, const double* tau, double* cumsum, int32_t* best_idx) {
  RG_API_BEGIN
  RG_CHECK(false, "rg_l1_fit: not implemented yet");
  RG_API_END
}

int rg_loco(rg_handle h, const int32_t* chr_of_block, double* pred_out) {
  RG_API_BEGIN
  RG_CHECK(false, "rg_loco: not implemented yet");
  RG_API_END
}

int rg_step2_create(const rg_step2_config* cfg, const double* X, const uint8_t* mask,
                    const uint8_t* in_analysis, rg_handle* out) {
  RG_API_BEGIN
  RG_CHECK(false, "rg_step2_create: not implemented yet");
  RG_API_END
}

int rg_s2_set_chr(rg_handle h, const double* res) {
  RG_API_BEGIN
  RG_CHECK(false, "rg_s2_set_chr: not implemented yet");
  RG_API_END
}

int rg_s2_block_bed(rg_handle h, const uint8_t* packed, int64_t row_stride, int32_t bs,
                    const int32_t* sample_idx, int32_t ref_first, const rg_s2_out* out) {
  RG_API_BEGIN
  RG_CHECK(false, "rg_s2_block_bed: not implemented yet");
  RG_API_END
}

int rg_W_info(rg_handle h, int32_t ph, void** dev_ptr, int64_t* ld, int64_t* ncols) {
  RG_API_BEGIN
  RG_CHECK(h && h->kind == 1 && ph >= 0 && ph < h->P, "bad argument");
  if (dev_ptr) *dev_ptr = h->W.p + (size_t)ph * h->Npad * h->B;
  if (ld) *ld = h->Npad;
  if (ncols) *ncols = h->B;
  RG_API_END
}

}  // extern "C"
