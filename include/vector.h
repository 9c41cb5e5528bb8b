/* Forwarding header: the reference's src/vector.h declarations live in sslsim.h. */
#include "sslsim.h"
