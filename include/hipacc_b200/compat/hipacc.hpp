// compat/hipacc.hpp -- lets an unmodified Hipacc DSL source (`#include "hipacc.hpp"`, the reference's dsl/hipacc.hpp) build
// against the B200 front: add -I<repo>/include/hipacc_b200/compat and compile the file with nvcc (-x cu) so that its
// kernel() bodies become device code (see ../hipacc.hpp).
#ifndef HIPACC_B200_COMPAT_HIPACC_HPP
#define HIPACC_B200_COMPAT_HIPACC_HPP
#include "../hipacc.hpp"
#endif
