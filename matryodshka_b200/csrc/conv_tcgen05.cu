// placeholder until the tcgen05 kernel lands
#include "net_internal.cuh"
namespace msi {
int conv_tc_plan_create(LayerPlan&, const ActBuf*, int, int) { set_error("tcgen05 back end not built yet"); return MSI_ERR_UNSUPPORTED; }
void conv_tc_plan_destroy(LayerPlan&) {}
int conv_tc_forward(const LayerPlan&, int, float*, cudaStream_t) { return MSI_ERR_UNSUPPORTED; }
int conv_tc_pack_weights(LayerPlan&, const ActBuf*, cudaStream_t) { return MSI_ERR_UNSUPPORTED; }
}
