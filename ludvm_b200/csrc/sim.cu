// sim.cu -- placeholder entry points (replaced by the on-device time loop).
#include "common.cuh"
using namespace ludvm;
LUDVM_API int ludvm_sim_create(ludvm_ctx *, const ludvm_sim_params *, const ludvm_sim_tables *, ludvm_sim **)
{ return set_error(LUDVM_E_UNSUPPORTED, "time loop not built yet"); }
LUDVM_API int ludvm_sim_run(ludvm_sim *, long) { return set_error(LUDVM_E_UNSUPPORTED, "time loop not built yet"); }
LUDVM_API int ludvm_sim_steps_done(ludvm_sim *, long *) { return set_error(LUDVM_E_UNSUPPORTED, "time loop not built yet"); }
LUDVM_API int ludvm_sim_fetch(ludvm_sim *, int, void *, size_t) { return set_error(LUDVM_E_UNSUPPORTED, "time loop not built yet"); }
LUDVM_API int ludvm_sim_field_bytes(ludvm_sim *, int, size_t *) { return set_error(LUDVM_E_UNSUPPORTED, "time loop not built yet"); }
LUDVM_API int ludvm_sim_destroy(ludvm_sim *) { return LUDVM_OK; }
