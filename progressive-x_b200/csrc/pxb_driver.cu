// pxb_driver.cu -- task-level entry points (placeholder).
#include "pxb_internal.h"
using namespace pxb;
extern "C" {
int pxb_find_homographies(pxb_ctx *, const double *, int64_t, int64_t *, double *, int64_t, size_t, size_t, size_t,
                          size_t, double, double, double, double, double, size_t, size_t, int, size_t, double, int,
                          uint64_t) {
	set_error("not implemented yet");
	return PXB_ERR_UNSUPPORTED;
}
int pxb_find_two_view_motions(pxb_ctx *, const double *, int64_t, int64_t *, double *, int64_t, size_t, size_t,
                              size_t, size_t, double, double, double, double, double, size_t, size_t, int, size_t,
                              double, int, uint64_t) {
	set_error("not implemented yet");
	return PXB_ERR_UNSUPPORTED;
}
int pxb_find_6d_poses(pxb_ctx *, const double *, const double *, const double *, int64_t, int64_t *, double *,
                      int64_t, double, double, double, double, double, size_t, size_t, int, uint64_t) {
	set_error("not implemented yet");
	return PXB_ERR_UNSUPPORTED;
}
}
