// ca_l1_inst.cu -- one (NS, BEAMS) variant of k_lindblad1 per translation unit (-DCA_INST_NS=.. -DCA_INST_BEAMS=..), with and
// without the sweep prologue, so that the seven variants compile in parallel (see Makefile).
#include "ca_lindblad1.cuh"

#ifndef CA_INST_NS
#error "compile with -DCA_INST_NS=9|15|21 -DCA_INST_BEAMS=0|1|2|3"
#endif

namespace ca {
#define CA_L1_NAME2(NS, BEAMS) launch_l1_##NS##_##BEAMS
#define CA_L1_NAME(NS, BEAMS) CA_L1_NAME2(NS, BEAMS)
cudaError_t CA_L1_NAME(CA_INST_NS, CA_INST_BEAMS)(bool sweep, cudaStream_t stream, const R1Params& P, int grid) {
    return sweep ? launch_lindblad1<CA_INST_NS, true, CA_INST_BEAMS>(stream, P, grid)
                 : launch_lindblad1<CA_INST_NS, false, CA_INST_BEAMS>(stream, P, grid);
}
}  // namespace ca
