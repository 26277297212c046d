// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace loosb200 {
bool tensor_path_available() { return false; }
int launch_pairs_tensor(const PairJob&, cudaStream_t, int*) { return LOOSB200_ERR_UNSUPPORTED; }
}
