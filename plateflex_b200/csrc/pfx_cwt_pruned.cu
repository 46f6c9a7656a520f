// pfx_cwt_pruned.cu -- band-pruned inverse transform engine (placeholder until implemented).
#include "pfx_plan.hpp"

namespace pfx {

int plan_init_pruned(pfx_plan *) { return PFX_OK; }

int pruned_scale_planes(pfx_plan *, int, int)
{
    set_error("pruned engine not implemented yet");
    return PFX_ERR_STATE;
}

PlaneView pruned_view(const pfx_plan *, int) { return PlaneView{}; }

}  // namespace pfx
