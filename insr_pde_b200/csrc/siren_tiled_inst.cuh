// siren_tiled_inst.cuh -- instantiation + dispatch of the tiled family for every (D, O, order).
#define INSR_TILED_CASE(D_, O_, ORD_, CALL) \
    if (dm.D == D_ && dm.O == O_ && order == ORD_) { constexpr int D = D_, O = O_, ORDER = ORD_; \
        (void)D; (void)O; (void)ORDER; return CALL; }
#define INSR_TILED_O(D_, ORD_, CALL) INSR_TILED_CASE(D_, 1, ORD_, CALL) INSR_TILED_CASE(D_, 2, ORD_, CALL) INSR_TILED_CASE(D_, 3, ORD_, CALL)
#define INSR_TILED_ORD(D_, CALL) INSR_TILED_O(D_, 0, CALL) INSR_TILED_O(D_, 1, CALL) INSR_TILED_O(D_, 2, CALL) INSR_TILED_O(D_, 3, CALL)
#define INSR_TILED_ALL(CALL) INSR_TILED_ORD(1, CALL) INSR_TILED_ORD(2, CALL) INSR_TILED_ORD(3, CALL)

int insr_tiled_forward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N, float *y,
                       float *jac, float *h2, float *ws, void *stream, int64_t *launches, bool tensor, bool keep_tape) {
    INSR_TILED_ALL((insr_tiled::run_forward<D, O, ORDER>(dm, theta, x, N, y, jac, h2, ws, stream, launches, tensor, keep_tape)))
    return -6;
}
int insr_tiled_backward(const SirenDims &dm, int order, const float *theta, const float *x, int64_t N,
                        const float *gy, const float *gjac, const float *gh2, float *gtheta, float *gx, float *ws,
                        void *stream, int64_t *launches, bool tensor, bool have_tape) {
    INSR_TILED_ALL((insr_tiled::run_backward<D, O, ORDER>(dm, theta, x, N, gy, gjac, gh2, gtheta, gx, ws, stream, launches, tensor, have_tape)))
    return -6;
}

#if defined(INSR_WIDE_PROFILE) && !defined(INSR_CPU_EMU)
// debug builds only (tools/wide_phase_profile.py): read / reset the phase counters of k_wide_tc
extern "C" int insr_debug_wide_prof(unsigned long long *out16, int reset) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out16, insr_wide::g_wide_prof, 16 * sizeof(unsigned long long));
    if (reset) { unsigned long long z[16] = {0}; cudaMemcpyToSymbol(insr_wide::g_wide_prof, z, sizeof(z)); }
    return 0;
}
#endif
