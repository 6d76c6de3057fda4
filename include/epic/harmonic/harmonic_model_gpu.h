/* Forwarding header: the reference callers include <epic/harmonic/harmonic_model_gpu.h>; all declarations live in <epic/libepic.h>. */
#include "../libepic.h"
