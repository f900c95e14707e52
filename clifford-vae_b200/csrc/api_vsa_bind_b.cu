// bind-family launchers, part b (see bind_launch.cuh)
#include "bind_launch.cuh"
#include "../../include/clifford_b200.h"

using namespace cvb;

extern "C" int cvb_internal_bind_b(const BindParams* p, int d, int mode, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  switch (mode) {
    case CVB_BIND_DIV: return dispatch_bind<kBindDiv>(*p, d, st);
    case CVB_BIND_DIV_CONJ: return dispatch_bind<kBindDivConj>(*p, d, st);
    case CVB_BIND_NEG_MUL_CONJ: return dispatch_bind<kBindNegMulConj>(*p, d, st);
  }
  return kBadArgument;
}
