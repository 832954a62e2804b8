// placeholder until the tcgen05 engine lands
#include "common.cuh"
namespace nvqa {
struct UmmaWorkspace { int dummy; };
int umma_gemm(cudaStream_t, int, bool, bool, int, int, int, const float*, int, const float*, int, float*, int, bool,
              const float*, const float*, UmmaWorkspace*) { set_error("tcgen05 GEMM engine not built"); return 1; }
int umma_workspace_create(UmmaWorkspace** ws, size_t) { *ws = new UmmaWorkspace(); return 0; }
void umma_workspace_destroy(UmmaWorkspace* ws) { delete ws; }
}
