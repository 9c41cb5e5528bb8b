/* Forwarding header: the reference's src/team.h declarations live in sslsim.h. */
#include "sslsim.h"
