/* Forwarding header: the reference callers include <epic/constants.h>; the definitions live in <epic/libepic.h>. */
#include "libepic.h"
