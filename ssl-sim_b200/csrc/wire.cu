/*
 * wire.cu -- proto2 wire encoding/decoding for the ssl-vision detection/geometry packets and the grSim
 * command/replacement packets (host code only; lives in a .cu file just to share the build rule).
 * Replaces the protobuf-generated code the reference links (reference src/serialize.cpp:10-13,
 * CMakeLists.txt:35-45); field numbers cite the .proto files of the reference.
 */
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>
#include "batch_internal.h"
#include "serialize.h"

namespace {

struct Writer {
  unsigned char *p, *end;
  bool ok;
  Writer(void *buf, size_t cap) : p((unsigned char *)buf), end((unsigned char *)buf + cap), ok(true) {}
  void byte(unsigned v) { if (p < end) *p++ = (unsigned char)v; else ok = false; }
  void varint(uint64_t v) { while (v >= 0x80) { byte((unsigned)(v & 0x7F) | 0x80); v >>= 7; } byte((unsigned)v); }
  void tag(int field, int wt) { varint((uint64_t)((field << 3) | wt)); }
  void f32(int field, float v) { tag(field, 5); uint32_t b; memcpy(&b, &v, 4); for (int i = 0; i < 4; ++i) byte((b >> (8 * i)) & 0xFF); }
  void f64(int field, double v) { tag(field, 1); uint64_t b; memcpy(&b, &v, 8); for (int i = 0; i < 8; ++i) byte((unsigned)((b >> (8 * i)) & 0xFF)); }
  void u32(int field, uint32_t v) { tag(field, 0); varint(v); }
  void i32(int field, int32_t v) { tag(field, 0); varint((uint64_t)(int64_t)v); } /* negative: sign-extended, 10 bytes */
  void bytes(int field, const void *d, size_t n) {
    tag(field, 2); varint(n);
    if ((size_t)(end - p) >= n) { memcpy(p, d, n); p += n; } else ok = false;
  }
};

/* SSL_DetectionBall: protos/messages_robocup_ssl_detection.proto:1-9 */
size_t encode_ball(const float *r, unsigned char *buf, size_t cap) {
  Writer w(buf, cap);
  w.f32(1, r[0]); w.f32(3, r[1]); w.f32(4, r[2]); w.f32(5, r[3]); w.f32(6, r[4]); w.f32(7, r[5]);
  return w.ok ? (size_t)(w.p - buf) : 0;
}
/* SSL_DetectionRobot: protos/messages_robocup_ssl_detection.proto:11-20 */
size_t encode_robot(const float *r, unsigned char *buf, size_t cap) {
  Writer w(buf, cap);
  w.f32(1, r[0]); w.u32(2, (uint32_t)r[1]); w.f32(3, r[2]); w.f32(4, r[3]); w.f32(5, r[4]); w.f32(6, r[5]); w.f32(7, r[6]);
  return w.ok ? (size_t)(w.p - buf) : 0;
}

struct Reader {
  const unsigned char *p, *end;
  bool ok;
  Reader(const void *d, size_t n) : p((const unsigned char *)d), end((const unsigned char *)d + n), ok(true) {}
  bool more() const { return ok && p < end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int s = 0; s < 70; s += 7) {
      if (p >= end) { ok = false; return 0; }
      const unsigned b = *p++;
      v |= (uint64_t)(b & 0x7F) << s;
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  float f32() { if (end - p < 4) { ok = false; return 0; } uint32_t b = 0; for (int i = 0; i < 4; ++i) b |= (uint32_t)p[i] << (8 * i); p += 4; float v; memcpy(&v, &b, 4); return v; }
  double f64() { if (end - p < 8) { ok = false; return 0; } uint64_t b = 0; for (int i = 0; i < 8; ++i) b |= (uint64_t)p[i] << (8 * i); p += 8; double v; memcpy(&v, &b, 8); return v; }
  Reader sub() { const uint64_t n = varint(); if (!ok || (uint64_t)(end - p) < n) { ok = false; return Reader(p, 0); } Reader r(p, (size_t)n); p += n; return r; }
  void skip(int wt) {
    if (wt == 0) varint(); else if (wt == 1) { if (end - p < 8) ok = false; else p += 8; }
    else if (wt == 2) sub(); else if (wt == 5) { if (end - p < 4) ok = false; else p += 4; } else ok = false;
  }
};

} // namespace

extern "C" {

/* SSL_WrapperPacket{detection = 1} / SSL_DetectionFrame: protos/messages_robocup_ssl_wrapper.proto:4-7,
 * messages_robocup_ssl_detection.proto:22-30.  protobuf writes fields in field-number order:
 * frame_number 1, t_capture 2, t_sent 3, camera_id 4, balls 5, robots_yellow 6, robots_blue 7. */
int sslsim_encode_detection(unsigned int frame_number, double t_capture, double t_sent, const float *record,
                            const int *teams, int n_robots, char *buffer, int buffer_size) {
  if (!record || !buffer || buffer_size < 0 || n_robots < 0) return -1;
  std::vector<unsigned char> frame((size_t)64 + 40 + (size_t)n_robots * 48);
  Writer f(frame.data(), frame.size());
  f.u32(1, frame_number);
  f.f64(2, t_capture);
  f.f64(3, t_sent);
  f.u32(4, 0); /* camera_id: set explicitly upstream (src/serialize.cpp:23), so it is on the wire */
  unsigned char tmp[64];
  f.bytes(5, tmp, encode_ball(record, tmp, sizeof tmp));
  for (int pass = 0; pass < 2; ++pass) { /* yellow list (6) first, then blue (7), each in world order */
    const int team = pass == 0 ? TEAM_YELLOW : TEAM_BLUE;
    for (int i = 0; i < n_robots; ++i)
      if (teams[i] == team) f.bytes(pass == 0 ? 6 : 7, tmp, encode_robot(record + 6 + 7 * i, tmp, sizeof tmp));
  }
  if (!f.ok) return -1;
  Writer w(buffer, (size_t)buffer_size);
  w.bytes(1, frame.data(), (size_t)(f.p - frame.data()));
  return w.ok ? (int)(w.p - (unsigned char *)buffer) : -1;
}

/* reference src/serialize.cpp:17-76 */
int serialize_world(const struct World *world, char *buffer, int buffer_size) {
  if (!world) return -1;
  const int R = world_robot_count(world);
  std::vector<float> rec((size_t)6 + 7 * (size_t)(R > 0 ? R : 0));
  std::vector<int> teams((size_t)(R > 0 ? R : 1));
  if (world_ball_count(world) < 1) return -1;
  const struct Vec3 b = ball_get_vec(world_get_ball(world, 0));
  rec[0] = 1.0f; rec[1] = b.x; rec[2] = b.y; rec[3] = b.z; rec[4] = b.x; rec[5] = b.y;
  for (int i = 0; i < R; ++i) {
    const struct Robot *r = world_get_robot(world, i);
    const struct Pos2 p = robot_get_pos(r);
    float *o = &rec[6 + 7 * (size_t)i];
    o[0] = 1.0f; o[1] = (float)get_id(r); o[2] = p.x; o[3] = p.y; o[4] = p.w; o[5] = p.x; o[6] = p.y;
    teams[i] = (int)get_team(r);
  }
  const double t_sent = std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count();
  return sslsim_encode_detection(world_get_frame_number(world), world_get_timestamp(world), t_sent, rec.data(),
                                 teams.data(), R, buffer, buffer_size);
}

/* reference src/serialize.cpp:78-104: float metres stored into the int32 fields (C++ implicit conversion,
 * i.e. truncation); SSL_WrapperPacket{geometry = 2 {field = 1}} */
int serialize_field(const struct FieldGeometry *field, char *buffer, int buffer_size) {
  if (!field || !buffer || buffer_size < 0) return -1;
  unsigned char fs[14 * 11], gd[14 * 11 + 8];
  Writer f(fs, sizeof fs);
  const float v[14] = {field->line_width, field->field_length, field->field_width, field->boundary_width,
                       field->referee_width, field->goal_width, field->goal_depth, field->goal_wall_width,
                       field->center_circle_radius, field->defense_radius, field->defense_stretch,
                       field->free_kick_from_defense_dist, field->penalty_spot_from_field_line_dist,
                       field->penalty_line_from_spot_dist};
  for (int k = 0; k < 14; ++k) f.i32(k + 1, (int32_t)v[k]);
  Writer g(gd, sizeof gd);
  g.bytes(1, fs, (size_t)(f.p - fs));
  Writer w(buffer, (size_t)buffer_size);
  w.bytes(2, gd, (size_t)(g.p - gd));
  return (f.ok && g.ok && w.ok) ? (int)(w.p - (unsigned char *)buffer) : -1;
}

int world_batch_serialize_detections(struct WorldBatch *b, const int *world_ids, int n, char *out, size_t capacity,
                                     int *sizes) {
  if (!b || n < 0 || (n > 0 && (!world_ids || !out))) return SSLSIM_E_ARG;
  const size_t rec = world_batch_detection_record_size(b);
  std::vector<char> recs((size_t)n * rec + 1);
  const int rc = world_batch_gather_detections(b, world_ids, n, recs.data(), recs.size());
  if (rc < 0) return rc;
  const double t_sent = std::chrono::duration<double>(std::chrono::high_resolution_clock::now().time_since_epoch()).count();
  size_t off = 0;
  for (int i = 0; i < n; ++i) {
    const size_t room = capacity - off;
    const int sz = sslsim_encode_detection(b->frame, (double)b->timestamp, t_sent, (const float *)(recs.data() + (size_t)i * rec),
                                           b->robot_teams.data(), b->R, out + off, room > 0x7fffffff ? 0x7fffffff : (int)room);
    if (sz < 0) return SSLSIM_E_SPACE;
    if (sizes) sizes[i] = sz;
    off += (size_t)sz;
  }
  return (int)off;
}

/* grSim_Packet{commands = 1, replacement = 2}: protos/grSim_Packet.proto:3-6 */
int sslsim_decode_grsim_packet(const void *data, size_t len, int n_robots, struct RobotCommand *cmds,
                               struct RobotReplacement *robots, int *n_robots_rep, struct BallReplacement *ball,
                               int *has_ball, double *timestamp) {
  if (!data && len) return -1;
  const int rep_cap = n_robots_rep ? *n_robots_rep : 0;
  if (n_robots_rep) *n_robots_rep = 0;
  if (has_ball) *has_ball = 0;
  int n_cmd = 0;
  Reader pk(data, len);
  while (pk.more()) {
    const uint64_t t = pk.varint();
    const int fld = (int)(t >> 3), wt = (int)(t & 7);
    if (fld == 1 && wt == 2) { /* grSim_Commands: protos/grSim_Commands.proto:16-20 */
      Reader c = pk.sub();
      bool yellow = false;
      std::vector<std::pair<unsigned, RobotCommand>> list;
      while (c.more()) {
        const uint64_t ct = c.varint();
        const int cf = (int)(ct >> 3), cw = (int)(ct & 7);
        if (cf == 1 && cw == 1) { const double ts = c.f64(); if (timestamp) *timestamp = ts; }
        else if (cf == 2 && cw == 0) yellow = c.varint() != 0;
        else if (cf == 3 && cw == 2) { /* grSim_Robot_Command: protos/grSim_Commands.proto:1-14 */
          Reader r = c.sub();
          unsigned id = 0;
          bool wheelsspeed = false, spinner = false;
          float vt = 0, vn = 0, va = 0, w1 = 0, w2 = 0, w3 = 0, w4 = 0, kx = 0, kz = 0;
          while (r.more()) {
            const uint64_t rt = r.varint();
            const int rf = (int)(rt >> 3), rw = (int)(rt & 7);
            if (rf == 1 && rw == 0) id = (unsigned)r.varint();
            else if (rf == 2 && rw == 5) kx = r.f32();
            else if (rf == 3 && rw == 5) kz = r.f32();
            else if (rf == 4 && rw == 5) vt = r.f32();
            else if (rf == 5 && rw == 5) vn = r.f32();
            else if (rf == 6 && rw == 5) va = r.f32();
            else if (rf == 7 && rw == 0) spinner = r.varint() != 0;
            else if (rf == 8 && rw == 0) wheelsspeed = r.varint() != 0;
            else if (rf == 9 && rw == 5) w1 = r.f32();
            else if (rf == 10 && rw == 5) w2 = r.f32();
            else if (rf == 11 && rw == 5) w3 = r.f32();
            else if (rf == 12 && rw == 5) w4 = r.f32();
            else r.skip(rw);
          }
          if (!r.ok) return -1;
          RobotCommand rc;
          memset(&rc, 0, sizeof rc);
          rc.flags = SSLSIM_CMD_DRIVE | (wheelsspeed ? SSLSIM_CMD_WHEELS : 0u) | (spinner ? SSLSIM_CMD_SPINNER : 0u);
          if (wheelsspeed) { rc.wheel[0] = w1; rc.wheel[1] = w2; rc.wheel[2] = w3; rc.wheel[3] = w4; }
          else { rc.wheel[0] = vt; rc.wheel[1] = vn; rc.wheel[2] = va; rc.wheel[3] = 0.0f; }
          rc.kickspeedx = kx; rc.kickspeedz = kz;
          list.push_back(std::make_pair(id, rc));
        } else c.skip(cw);
      }
      if (!c.ok) return -1;
      for (auto &e : list) { /* isteamyellow may follow the robot commands on the wire: apply afterwards */
        const long slot = 2L * (long)e.first + (yellow ? 1 : 0);
        if (cmds && slot < n_robots) { cmds[slot] = e.second; ++n_cmd; }
      }
    } else if (fld == 2 && wt == 2) { /* grSim_Replacement: protos/grSim_Replacement.proto:16-19 */
      Reader rp = pk.sub();
      while (rp.more()) {
        const uint64_t rt = rp.varint();
        const int rf = (int)(rt >> 3), rw = (int)(rt & 7);
        if (rf == 1 && rw == 2) { /* grSim_BallReplacement :9-14 (doubles) */
          Reader bb = rp.sub();
          BallReplacement br;
          memset(&br, 0, sizeof br);
          while (bb.more()) {
            const uint64_t bt = bb.varint();
            const int bf = (int)(bt >> 3), bw = (int)(bt & 7);
            if (bw == 1 && bf >= 1 && bf <= 4) { const float v = (float)bb.f64(); (bf == 1 ? br.x : bf == 2 ? br.y : bf == 3 ? br.vx : br.vy) = v; }
            else bb.skip(bw);
          }
          if (!bb.ok) return -1;
          if (ball) *ball = br;
          if (has_ball) *has_ball = 1;
        } else if (rf == 2 && rw == 2) { /* grSim_RobotReplacement :1-7 */
          Reader rr = rp.sub();
          double x = 0, y = 0, dir = 0;
          unsigned id = 0;
          bool yellow = false;
          while (rr.more()) {
            const uint64_t bt = rr.varint();
            const int bf = (int)(bt >> 3), bw = (int)(bt & 7);
            if (bf == 1 && bw == 1) x = rr.f64();
            else if (bf == 2 && bw == 1) y = rr.f64();
            else if (bf == 3 && bw == 1) dir = rr.f64();
            else if (bf == 4 && bw == 0) id = (unsigned)rr.varint();
            else if (bf == 5 && bw == 0) yellow = rr.varint() != 0;
            else rr.skip(bw);
          }
          if (!rr.ok) return -1;
          const long slot = 2L * (long)id + (yellow ? 1 : 0);
          if (robots && n_robots_rep && *n_robots_rep < rep_cap && slot < n_robots) {
            RobotReplacement &o = robots[(*n_robots_rep)++];
            o.world = 0; o.robot = (int32_t)slot; o.x = (float)x; o.y = (float)y; o.dir = (float)dir;
          }
        } else rp.skip(rw);
      }
      if (!rp.ok) return -1;
    } else pk.skip(wt);
  }
  return pk.ok ? n_cmd : -1;
}

} /* extern "C" */
