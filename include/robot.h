/* Forwarding header: the reference's src/robot.h declarations live in sslsim.h. */
#include "sslsim.h"
