// softening kernel id 2: default tuning variant only
#include "nb_dispatch.cuh"
NB_DEFINE_KID(2, false)
