// placeholder until the tcgen05 kernel lands (next commit)
#include "common.cuh"
extern "C" {
size_t jaf_convlstm_wpack_bytes(int Cin, int Ch) { return (size_t)9 * 4 * Ch * (Cin + Ch) * 2; }
int jaf_convlstm_pack_weight(const float*, int, int, void*, void*) {
  jaf::set_error("jaf_convlstm_pack_weight: tensor-core path not built");
  return JAF_ERR_UNSUPPORTED;
}
int jaf_convlstm_step_tc(const void*, const void*, const float*, const void*, const float*, int, int, int, int, int,
                         void*, float*, void*) {
  jaf::set_error("jaf_convlstm_step_tc: tensor-core path not built");
  return JAF_ERR_UNSUPPORTED;
}
}
