/*
 * ingest.cu -- include/sslsim_ingest.h: UDP receive -> grSim_Packet decode -> device command rows / K2 for the
 * steered worlds.  The receiver is set up like the reference's (src/utils/net.c:72-93: SO_REUSEADDR, bind,
 * IP_ADD_MEMBERSHIP; :135-146: recvfrom), written here from scratch and non-blocking.
 */
#include <arpa/inet.h>
#include <fcntl.h>
#include <netinet/in.h>
#include <netinet/ip.h>
#include <poll.h>
#include <sys/socket.h>
#include <unistd.h>
#include <cerrno>
#include <cstdio>
#include <cstring>
#include <new>
#include <vector>
#include "batch_internal.h"
#include "sslsim_ingest.h"

struct GrSimIngest {
  WorldBatch *b = nullptr;
  int base_port = 0;
  std::vector<int> world_ids, fds;
  RobotCommand *h_cmd = nullptr;          /* pinned: [n][R], the current command row of every steered world */
  std::vector<unsigned char> touched;
  std::vector<RobotReplacement> rrep;
  std::vector<BallReplacement> brep;
  std::vector<double> last_ts;
  unsigned long long packets = 0, bad = 0;
};

static bool is_multicast(in_addr_t a) { return (ntohl(a) >> 28) == 0xE; } /* 224.0.0.0/4 */

extern "C" {

void delete_grsim_ingest(struct GrSimIngest *g) {
  if (!g) return;
  for (int fd : g->fds) if (fd >= 0) close(fd);
  if (g->h_cmd) { cudaSetDevice(g->b->device); cudaStreamSynchronize(g->b->stream); cudaFreeHost(g->h_cmd); }
  delete g;
}

struct GrSimIngest *new_grsim_ingest(struct WorldBatch *b, const int *world_ids, int n, int base_port, const char *addr,
                                     const char *iface) {
  if (!b || !world_ids || n <= 0 || n > 4096 || base_port <= 0 || base_port + n > 65536) return nullptr;
  for (int k = 0; k < n; ++k) if (world_ids[k] < 0 || world_ids[k] >= b->W) return nullptr;
  GrSimIngest *g = new (std::nothrow) GrSimIngest();
  if (!g) return nullptr;
  g->b = b; g->base_port = base_port;
  g->world_ids.assign(world_ids, world_ids + n);
  g->fds.assign(n, -1); g->touched.assign(n, 0); g->last_ts.assign(n, 0.0);
  cudaSetDevice(b->device);
  const size_t rows = (size_t)n * (b->R > 0 ? b->R : 1);
  if (cudaMallocHost(&g->h_cmd, rows * sizeof(RobotCommand)) != cudaSuccess) { g->h_cmd = nullptr; delete_grsim_ingest(g); return nullptr; }
  memset(g->h_cmd, 0, rows * sizeof(RobotCommand)); /* passive robots until a packet says otherwise */
  const in_addr_t group = (addr && *addr) ? inet_addr(addr) : htonl(INADDR_ANY);
  const in_addr_t ifc = (iface && *iface) ? inet_addr(iface) : htonl(INADDR_ANY);
  for (int k = 0; k < n; ++k) {
    const int fd = socket(PF_INET, SOCK_DGRAM, IPPROTO_IP);
    if (fd < 0) { perror("Could not open socket file descriptor"); delete_grsim_ingest(g); return nullptr; }
    g->fds[k] = fd;
    const unsigned int one = 1;
    if (setsockopt(fd, SOL_SOCKET, SO_REUSEADDR, &one, sizeof one) < 0) { perror("Could not enable address reuse"); delete_grsim_ingest(g); return nullptr; }
    sockaddr_in sa;
    memset(&sa, 0, sizeof sa);
    sa.sin_family = PF_INET; sa.sin_port = htons((unsigned short)(base_port + k)); sa.sin_addr.s_addr = group;
    if (bind(fd, (sockaddr *)&sa, sizeof sa) < 0) { perror("Could not bind socket to interface"); delete_grsim_ingest(g); return nullptr; }
    if (is_multicast(group)) {
      ip_mreq mreq;
      mreq.imr_multiaddr.s_addr = group; mreq.imr_interface.s_addr = ifc;
      if (setsockopt(fd, IPPROTO_IP, IP_ADD_MEMBERSHIP, &mreq, sizeof mreq) < 0) { perror("Could not join to multicast"); delete_grsim_ingest(g); return nullptr; }
    }
    const int fl = fcntl(fd, F_GETFL, 0);
    if (fl < 0 || fcntl(fd, F_SETFL, fl | O_NONBLOCK) < 0) { perror("Could not make the socket non-blocking"); delete_grsim_ingest(g); return nullptr; }
  }
  return g;
}

int grsim_ingest_poll(struct GrSimIngest *g, int timeout_ms) {
  SSLSIM_RANGE("grsim_ingest_poll");
  if (!g) return SSLSIM_E_ARG;
  const int n = (int)g->fds.size(), R = g->b->R;
  std::vector<pollfd> pf(n);
  for (int k = 0; k < n; ++k) { pf[k].fd = g->fds[k]; pf[k].events = POLLIN; pf[k].revents = 0; }
  const int rdy = poll(pf.data(), n, timeout_ms);
  if (rdy < 0) return errno == EINTR ? 0 : SSLSIM_E_ARG;
  int consumed = 0;
  char buf[65536];
  std::vector<RobotReplacement> rr((size_t)(R > 0 ? R : 1));
  for (int k = 0; k < n; ++k) {
    if (!(pf[k].revents & POLLIN)) continue;
    for (;;) {
      const ssize_t len = recvfrom(g->fds[k], buf, sizeof buf, 0, nullptr, nullptr);
      if (len < 0) break; /* EAGAIN: drained */
      int nrr = (int)rr.size(), has_ball = 0;
      BallReplacement br;
      double ts = 0.0;
      const int rc = sslsim_decode_grsim_packet(buf, (size_t)len, R, g->h_cmd + (size_t)k * R, rr.data(), &nrr, &br, &has_ball, &ts);
      ++g->packets; ++consumed;
      if (rc < 0) { ++g->bad; continue; }
      g->touched[k] = 1;
      g->last_ts[k] = ts;
      for (int i = 0; i < nrr; ++i) { rr[i].world = g->world_ids[k]; g->rrep.push_back(rr[i]); }
      if (has_ball) { br.world = g->world_ids[k]; g->brep.push_back(br); }
    }
  }
  return consumed;
}

int grsim_ingest_apply(struct GrSimIngest *g) {
  SSLSIM_RANGE("grsim_ingest_apply");
  if (!g) return SSLSIM_E_ARG;
  WorldBatch *b = g->b;
  int rc = world_batch_sync(b); /* joins chunk streams; the steps queued so far read the old rows */
  if (rc < 0) return rc;
  cudaSetDevice(b->device);
  const int R = b->R;
  int updated = 0;
  bool any_cmd = false;
  for (size_t k = 0; k < g->touched.size(); ++k) any_cmd = any_cmd || g->touched[k];
  if (any_cmd && R > 0) {
    rc = sslsim_ensure_cmd_buffer(b);
    if (rc < 0) return rc;
    if (b->cmd_active && b->cmd_active != b->d_cmd) {
      /* the caller had installed its own device buffer: carry it over so untouched worlds keep their commands */
      if (!sslsim_check(b, cudaMemcpyAsync(b->d_cmd, b->cmd_active, (size_t)b->W * R * sizeof(RobotCommand), cudaMemcpyDeviceToDevice, b->stream), "ingest carry-over")) return SSLSIM_E_CUDA;
    }
    for (size_t k = 0; k < g->touched.size(); ++k) {
      if (!g->touched[k]) continue;
      if (!sslsim_check(b, cudaMemcpyAsync(b->d_cmd + (size_t)g->world_ids[k] * R, g->h_cmd + k * R, (size_t)R * sizeof(RobotCommand),
                                           cudaMemcpyHostToDevice, b->stream), "ingest H2D")) return SSLSIM_E_CUDA;
      g->touched[k] = 0;
      ++updated;
    }
    b->cmd_active = b->d_cmd;
    if (!sslsim_check(b, cudaStreamSynchronize(b->stream), "ingest sync")) return SSLSIM_E_CUDA; /* staging rows are reused */
  }
  if (!g->rrep.empty()) {
    rc = world_batch_replace_robots(b, g->rrep.data(), (int)g->rrep.size());
    g->rrep.clear();
    if (rc < 0) return rc;
  }
  if (!g->brep.empty()) {
    rc = world_batch_replace_balls(b, g->brep.data(), (int)g->brep.size());
    g->brep.clear();
    if (rc < 0) return rc;
  }
  return updated;
}

int grsim_ingest_port(const struct GrSimIngest *g, int k) { return (g && k >= 0 && k < (int)g->fds.size()) ? g->base_port + k : -1; }
unsigned long long grsim_ingest_packets(const struct GrSimIngest *g) { return g ? g->packets : 0; }
unsigned long long grsim_ingest_bad_packets(const struct GrSimIngest *g) { return g ? g->bad : 0; }
double grsim_ingest_last_timestamp(const struct GrSimIngest *g, int k) { return (g && k >= 0 && k < (int)g->last_ts.size()) ? g->last_ts[k] : 0.0; }

} /* extern "C" */
