// pxb_expansion.cu -- alpha-expansion label sweep (row a11). Placeholder until the GPU max-flow lands.
#include "pxb_internal.h"

namespace pxb {
int launch_alpha_expansion(pxb_ctx *, const double *, int64_t, int32_t, double, double, const int32_t *,
                           const int32_t *, int64_t, const int32_t *, int32_t *, double *) {
	set_error("alpha-expansion (lambda > 0) is not implemented yet");
	return PXB_ERR_UNSUPPORTED;
}
} // namespace pxb

extern "C" int pxb_lo_graph_cut(pxb_ctx *, const double *, const double *, const double *, int64_t, double,
                                const int32_t *, const int32_t *, uint8_t *) {
	pxb::set_error("graph-cut local optimisation with lambda > 0 is not implemented yet");
	return PXB_ERR_UNSUPPORTED;
}
