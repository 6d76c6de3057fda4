/* Forwarding header: the reference callers include <epic/harmonic/harmonic.h>; all declarations live in <epic/libepic.h>. */
#include "../libepic.h"
