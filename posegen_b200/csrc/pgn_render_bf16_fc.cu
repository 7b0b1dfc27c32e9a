// Frame-code (Optcodes) instantiations of the fused tcgen05 render kernel: the same source as pgn_render_bf16.cu compiled
// with kFC = true in its own translation unit (parallel build; the frame-code-free kernels carry none of the hooks).
#define PGN_RENDER_FC_UNIT 1
#include "pgn_render_bf16.cu"
