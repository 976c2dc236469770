// TEST INFRASTRUCTURE.  pybind glue (ours) that exposes only the four hash-grid
// embedding entry points of the UNMODIFIED reference translation units
// hashgrid/src/hashgrid_kernel.cu and hashgrid/src/hashgrid_bg_kernel.cu, so the
// parity tests can call the reference kernels without waiting for the ~40 min
// compile of hashgrid/src/rendering_kernel.cu that the reference's own
// hashgrid/binding.cpp drags in.  Prototypes come from the reference header
// hashgrid/include/hashgrid.h (included, not copied).
#include <pybind11/pybind11.h>
#include "hashgrid.h"

PYBIND11_MODULE(HASHGRID_EMBED, m) {
    m.def("embedding_forward_cuda", &embedding_forward_cuda, "");
    m.def("embedding_backward_cuda", &embedding_backward_cuda, "");
    m.def("embedding_bg_forward_cuda", &embedding_bg_forward_cuda, "");
    m.def("embedding_bg_backward_cuda", &embedding_bg_backward_cuda, "");
}
