// bind-family launchers, part a (see bind_launch.cuh)
#include "bind_launch.cuh"
#include "../../include/clifford_b200.h"

using namespace cvb;

extern "C" int cvb_internal_bind_a(const BindParams* p, int d, int mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case CVB_BIND_MUL: return dispatch_bind<kBindMul>(*p, d, st);
    case CVB_BIND_MUL_CONJ: return dispatch_bind<kBindMulConj>(*p, d, st);
  }
  return kBadArgument;
}
