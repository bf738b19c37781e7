// quotient kernel of one table (see quotient_kernel.cuh)
#include "quotient_kernel.cuh"
namespace zk { using namespace zkstark; ZK_INSTANTIATE_QUOTIENT(T_BYTE_PACKING) }
