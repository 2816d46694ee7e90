// softening kernel id 0: default tuning variant only
#include "nb_dispatch.cuh"
NB_DEFINE_KID(0, false)
