// softening kernel id 4 (plummer = roofline headline, spline = API default): all tuning variants
#include "nb_dispatch.cuh"
NB_DEFINE_KID(4, true)
