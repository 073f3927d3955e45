/* The whole compiled side of the package seen from R: the .Call shim of this repository.  Building the package from a
 * checkout compiles it in place; `tools/vendor.sh` copies the CUDA sources next to it for a self-contained tarball. */
#include "../../gpirt_b200/csrc/rshim/gpirt_rshim.c"
