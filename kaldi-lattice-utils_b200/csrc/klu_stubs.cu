// temporary stubs
#include "klu_common.cuh"
namespace klu {
int run_prune_dyn_beam(klu_ctx*, const klu_opts*) { set_error("prune-dyn-beam: not implemented"); return 1; }
int run_best_path2(klu_ctx*, const klu_opts*) { set_error("best-path2: not implemented"); return 1; }
int run_char_position(klu_ctx*, const klu_opts*) { set_error("char-position: not implemented"); return 1; }
}
using namespace klu;
extern "C" {
int klu_fetch_best_path2(klu_ctx*, int32_t*, float*, int32_t*) { set_error("not implemented"); return 1; }
int klu_fetch_prune(klu_ctx*, int32_t*, int32_t*, int32_t*, float*, float*, int32_t*, float*, float*, double*) { set_error("not implemented"); return 1; }
int klu_result_char_sizes(klu_ctx*, int64_t*) { set_error("not implemented"); return 1; }
int klu_fetch_char_position(klu_ctx*, int64_t*, int32_t*, int32_t*, int32_t*, int32_t*, double*) { set_error("not implemented"); return 1; }
int klu_topsort(int32_t, int64_t, int32_t*, int32_t*, int32_t*, int32_t*, float*, float*, float*, float*, int32_t*, int32_t*) { set_error("not implemented"); return 1; }
}
