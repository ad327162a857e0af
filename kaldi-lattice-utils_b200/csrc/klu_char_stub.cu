// temporary: char-position entry points (klu_char.cu replaces this file)
#include "klu_common.cuh"
namespace klu {
int run_char_position(klu_ctx*, const klu_opts*) { set_error("char-position: not implemented"); return 1; }
}
using namespace klu;
extern "C" {
int klu_result_char_sizes(klu_ctx*, int64_t*) { set_error("not implemented"); return 1; }
int klu_fetch_char_position(klu_ctx*, int64_t*, int32_t*, int32_t*, int32_t*, int32_t*, double*) { set_error("not implemented"); return 1; }
}
