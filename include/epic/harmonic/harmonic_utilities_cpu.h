/* Forwarding header: the reference callers include <epic/harmonic/harmonic_utilities_cpu.h>; all declarations live in <epic/libepic.h>. */
#include "../libepic.h"
