// pair-symmetric sweeps, softening kernel id 2 (see nb_sym_dispatch.cuh)
#include "nb_sym_dispatch.cuh"
NB_SYM_DEFINE_KID(2)
