/* Forwarding header: the reference's src/world.h declarations live in sslsim.h. */
#include "sslsim.h"
