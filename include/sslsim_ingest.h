/*
 * sslsim_ingest.h -- grSim command / replacement ingest for the worlds being steered.
 *
 * The reference compiles the grSim protos (CMakeLists.txt:43-45) and has a UDP receiver (src/utils/net.c:72-93
 * socket_receiver_bind, :135-146 socket_receive) but never wires them together: socket_receive has no caller and there
 * is no command input.  This is that missing loop, batched: UDP receive -> grSim_Packet decode
 * (protos/grSim_Packet.proto:3-6, grSim_Commands.proto, grSim_Replacement.proto; serialize.h
 * sslsim_decode_grsim_packet) -> device command rows and replacement scatter (K2) of the steered worlds only.
 *
 * One UDP port per steered world: world_ids[k] listens on base_port + k.  The socket set-up follows the reference's
 * receiver (address reuse, bind to the group address or INADDR_ANY, IP_ADD_MEMBERSHIP when the address is a multicast
 * group); a non-multicast address (e.g. "127.0.0.1") binds a plain unicast receiver.  Sockets are non-blocking:
 * grsim_ingest_poll waits at most timeout_ms for the first datagram and then drains everything that is queued.
 *
 * Host code only (POSIX sockets + the decoder); the device side is the batch's command buffer and K2.
 */
#ifndef SSLSIM_B200_INGEST_H
#define SSLSIM_B200_INGEST_H
#include "serialize.h"
#ifdef __cplusplus
extern "C" {
#endif

struct GrSimIngest;

/* addr "" = INADDR_ANY; iface "" = default interface (reference new_socket / new_socket_iface, src/utils/net.c:37-61).
 * NULL on failure (bad arguments, socket / bind / membership error: perror tells which, as the reference does). */
struct GrSimIngest *new_grsim_ingest(struct WorldBatch *batch, const int *world_ids, int n_worlds, int base_port,
                                     const char *addr, const char *iface);
void delete_grsim_ingest(struct GrSimIngest *ing);

/* receive + decode every queued datagram into the host-side staging rows; returns packets consumed (malformed ones
 * are counted by grsim_ingest_bad_packets and otherwise ignored) or <0 */
int grsim_ingest_poll(struct GrSimIngest *ing, int timeout_ms);
/* upload what the polls staged: command rows of the touched worlds into the batch's command buffer (which becomes the
 * active one; worlds that never received a packet keep passive robots), robot / ball replacements through K2 (re-dropped
 * from 0.5 m, ball woken: world_batch_replace_*).  Returns the number of worlds that were updated or <0. */
int grsim_ingest_apply(struct GrSimIngest *ing);

int grsim_ingest_port(const struct GrSimIngest *ing, int k);          /* base_port + k, -1 out of range */
unsigned long long grsim_ingest_packets(const struct GrSimIngest *ing);
unsigned long long grsim_ingest_bad_packets(const struct GrSimIngest *ing);
double grsim_ingest_last_timestamp(const struct GrSimIngest *ing, int k); /* grSim_Commands.timestamp of world k's last packet */

#ifdef __cplusplus
}
#endif
#endif
