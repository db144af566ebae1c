#include "krylov.cuh"
namespace fdfd {
int make_mg_precond(Helm2D&, const fdfd_krylov_opts&, Precond**, cudaStream_t) {
    set_error("multigrid preconditioner not built");
    return FDFD_ERR_STATE;
}
}
