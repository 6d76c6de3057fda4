/* Forwarding header: the reference callers include <epic/harmonic/harmonic_legacy_cpu.h>; all declarations live in <epic/libepic.h>. */
#include "../libepic.h"
