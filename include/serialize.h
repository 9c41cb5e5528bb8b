/*
 * serialize.h -- ssl-vision wire packets for the worlds being observed.
 *
 * serialize_world / serialize_field are ABI-identical to reference src/serialize.h:17-19 and produce the
 * bytes the reference's protobuf code produces (src/serialize.cpp:17-104; schemas
 * protos/messages_robocup_ssl_{wrapper,detection,geometry}.proto) with a small hand-rolled proto2 encoder:
 * there is no libprotobuf dependency.  Upstream quirks are kept on purpose and documented here:
 *   - positions are metres / radians, not ssl-vision millimetres (src/serialize.cpp:35,61 "TODO");
 *   - pixel_x / pixel_y repeat x / y;
 *   - serialize_field stores float metres into int32 fields, i.e. truncates (0.010 -> 0, 9.0 -> 9), and
 *     goal_height is not sent (src/serialize.cpp:82-98 vs messages_robocup_ssl_geometry.proto:2-15);
 *   - t_sent is the wall clock at encode time.
 * The sslsim_encode_* / sslsim_decode_* functions are pure host code (no CUDA) so they can be tested anywhere.
 */
#ifndef SSLSIM_B200_SERIALIZE_H
#define SSLSIM_B200_SERIALIZE_H
#define SERIALIZE_H /* the reference's guard */
#include <stddef.h>
#include "sslsim.h"
#include "sslsim_batch.h"
#ifdef __cplusplus
extern "C" {
#endif

/* -1 on error (buffer too small) or the size written: reference src/serialize.h:16-19 */
int serialize_world(const struct World *world, char *buffer, int buffer_size);
int serialize_field(const struct FieldGeometry *field, char *buffer, int buffer_size);

/* The encoder behind serialize_world: one SSL_WrapperPacket{detection} from a gathered detection record
 * (world_batch_gather_detections layout: 6 floats ball, then 7 floats per robot) and the roster teams
 * (enum Team per robot slot).  Deterministic: the caller supplies t_sent. */
int sslsim_encode_detection(unsigned int frame_number, double t_capture, double t_sent, const float *record,
                            const int *teams, int n_robots, char *buffer, int buffer_size);

/* n observed worlds -> n packets, concatenated in `out`; sizes[i] = bytes of packet i.
 * Returns the total number of bytes or <0 (SSLSIM_E_*). */
int world_batch_serialize_detections(struct WorldBatch *batch, const int *world_ids, int n, char *out,
                                     size_t capacity, int *sizes);

/* Input side: decode one grSim_Packet (protos/grSim_Packet.proto:3-6,
 * grSim_Commands.proto:1-20, grSim_Replacement.proto:1-19) for ONE world whose roster is B0,Y0,B1,Y1,...
 * (slot = 2 * id + isteamyellow, reference src/main.c:65-68).
 *   cmds     [n_robots] in/out: slots named by the packet are overwritten (SSLSIM_CMD_DRIVE set), others kept
 *   robots   out, capacity *n_robots_rep in; *n_robots_rep out = records written (.world left 0)
 *   ball     out; *has_ball = 1 if the packet carries a ball replacement
 * Returns the number of robot commands decoded, or <0 on a malformed packet. */
int sslsim_decode_grsim_packet(const void *data, size_t len, int n_robots, struct RobotCommand *cmds,
                               struct RobotReplacement *robots, int *n_robots_rep, struct BallReplacement *ball,
                               int *has_ball, double *timestamp);

#ifdef __cplusplus
}
#endif
#endif
