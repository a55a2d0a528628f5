// siren_fused_inst.cuh -- explicit instantiation of the fused kernels for ONE (D, O) pair.
// Included with INSR_INST_D / INSR_INST_O defined: once per pair by siren_fused_inst.cu (nvcc,
// one object per pair, built in parallel) or three times by insr_abi.cu in single-unit builds.
#define INSR_INST_CAT2(a, b, c) a##b##c
#define INSR_INST_CAT(a, b, c) INSR_INST_CAT2(a, b, c)

int INSR_INST_CAT(insr_fused_run_, INSR_INST_D, INSR_INST_O)(int kind, insr_fused::Params &p, int order,
                                                            void *stream, int64_t *launches) {
    using namespace insr_fused;
    constexpr int D = INSR_INST_D, O = INSR_INST_O;
#ifdef INSR_CPU_EMU
    if (kind == 3) kind = 0;             // the emulator has no tensor cores: FFMA kernels
    if (kind == 4) kind = 1;
    if (kind == 5) kind = 2;
#else
    if (kind == 3) {                     // tcgen05 / TMEM 3xTF32 forward (siren_tc.cuh)
        switch (order) {
            case 0: return insr_tc::launch_tc_fwd<D, O, 0>(p, stream, launches);
            case 1: return insr_tc::launch_tc_fwd<D, O, 1>(p, stream, launches);
            case 2: return insr_tc::launch_tc_fwd<D, O, 2>(p, stream, launches);
        }
        return -6;
    }
    if (kind == 4 || kind == 5) {        // tcgen05 backward / fused closure; p.ws carries the tape scratch
        float *ws = reinterpret_cast<float *>(p.ws);
        switch (order * 2 + (kind - 4)) {
            case 0: return insr_tc::launch_tc_bwd<D, O, 0, false>(p, ws, stream, launches);
            case 1: return insr_tc::launch_tc_bwd<D, O, 0, true>(p, ws, stream, launches);
            case 2: return insr_tc::launch_tc_bwd<D, O, 1, false>(p, ws, stream, launches);
            case 3: return insr_tc::launch_tc_bwd<D, O, 1, true>(p, ws, stream, launches);
            case 4: return insr_tc::launch_tc_bwd<D, O, 2, false>(p, ws, stream, launches);
            case 5: return insr_tc::launch_tc_bwd<D, O, 2, true>(p, ws, stream, launches);
        }
        return -6;
    }
#endif
    switch (order * 3 + kind) {
        case 0: return launch_fwd<D, O, 0>(p, stream, launches);
        case 1: return launch_bwd<D, O, 0, false>(p, stream, launches);
        case 2: return launch_bwd<D, O, 0, true>(p, stream, launches);
        case 3: return launch_fwd<D, O, 1>(p, stream, launches);
        case 4: return launch_bwd<D, O, 1, false>(p, stream, launches);
        case 5: return launch_bwd<D, O, 1, true>(p, stream, launches);
        case 6: return launch_fwd<D, O, 2>(p, stream, launches);
        case 7: return launch_bwd<D, O, 2, false>(p, stream, launches);
        case 8: return launch_bwd<D, O, 2, true>(p, stream, launches);
    }
    return -6;
}
#undef INSR_INST_D
#undef INSR_INST_O
