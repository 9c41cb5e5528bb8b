/* Forwarding header: the reference's src/ball.h declarations live in sslsim.h. */
#include "sslsim.h"
