// softening kernel id 3: default tuning variant only
#include "nb_dispatch.cuh"
NB_DEFINE_KID(3, false)
