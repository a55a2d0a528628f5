// translation unit of the fused mid-width tcgen05 kernels (siren_mid_tc.cuh)
#include "siren_mid_api.h"
#include "siren_mid_tc.cuh"

bool insr_mid_supported(const SirenDims &dm, int order) { return insr_mid::mid_supported(dm, order); }

#define INSR_MID_CASE(D_, O_, ORD_, CALL) \
    if (dm.D == D_ && dm.O == O_ && order == ORD_) { constexpr int D = D_, O = O_, ORDER = ORD_; \
        (void)D; (void)O; (void)ORDER; return CALL; }
#define INSR_MID_O(D_, ORD_, CALL) INSR_MID_CASE(D_, 1, ORD_, CALL) INSR_MID_CASE(D_, 2, ORD_, CALL) INSR_MID_CASE(D_, 3, ORD_, CALL)
// (D, order) pairs with at most 4 forward-mode streams
#define INSR_MID_ALL(CALL) \
    INSR_MID_O(1, 0, CALL) INSR_MID_O(1, 1, CALL) INSR_MID_O(1, 2, CALL) INSR_MID_O(1, 3, CALL) \
    INSR_MID_O(2, 0, CALL) INSR_MID_O(2, 1, CALL) INSR_MID_O(2, 2, CALL) INSR_MID_O(3, 0, CALL) INSR_MID_O(3, 1, CALL)

int insr_mid_width(const SirenDims &dm) { return insr_mid::hp16_of(dm.H) <= 64 ? 64 : 80; }

int insr_mid_forward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N, float *y, float *jac,
                     float *h2, float *Zpre, float *Act, size_t buf, void *stream, int64_t *launches) {
    if (!insr_mid::mid_supported(dm, order)) return -6;
    insr_mid::MidParams p{};
    p.dm = dm; p.theta = theta; p.x = x; p.N = N; p.y = y; p.jac = jac; p.h2 = h2;
    p.Zpre = Zpre; p.Act = Act; p.buf = (int64_t)buf;
    if (Zpre) {
        INSR_MID_ALL((insr_mid::launch_mid_fwd<D, O, ORDER, true>(p, stream, launches)))
    } else {
        INSR_MID_ALL((insr_mid::launch_mid_fwd<D, O, ORDER, false>(p, stream, launches)))
    }
    return -6;
}

int insr_mid_backward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N, const float *gy,
                      const float *gjac, const float *gh2, float *gtheta, float *gx, float *Zpre, float *Act, size_t buf,
                      void *stream, int64_t *launches) {
    if (!insr_mid::mid_supported(dm, order)) return -6;
    insr_mid::MidParams p{};
    p.dm = dm; p.theta = theta; p.x = x; p.N = N;
    p.Zpre = Zpre; p.Act = Act; p.buf = (int64_t)buf;
    p.gy = gy; p.gjac = gjac; p.gh2 = gh2; p.gx = gx; p.gtheta = gtheta;
    INSR_MID_ALL((insr_mid::launch_mid_bwd<D, O, ORDER>(p, stream, launches)))
    return -6;
}
