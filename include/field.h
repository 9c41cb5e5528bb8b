/* Forwarding header: the reference's src/field.h declarations live in sslsim.h. */
#include "sslsim.h"
