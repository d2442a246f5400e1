// libgfrender: tensor-core (tcgen05) field kernel -- placeholder until the probe-validated kernel lands.
#include "gf_model.cuh"

namespace gf {

int field_tc_launch(const GfModel* model, const FieldTcIO& io, cudaStream_t st) {
    (void)model; (void)io; (void)st;
    set_error("precision=1 (tcgen05 field) is not built in this revision");
    return GF_ERR_UNSUPPORTED;
}

}  // namespace gf
