/* Forwarding header: the reference callers include <epic/error_codes.h>; the definitions live in <epic/libepic.h>. */
#include "libepic.h"
