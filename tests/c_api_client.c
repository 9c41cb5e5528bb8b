/*
 * c_api_client.c -- a plain-C host program written against include/sslsim.h the way the reference's
 * CLI uses its API (reference src/main.c:62-96: new_world(&FIELD_2015), 6 blue + 6 yellow robots,
 * world_step in a loop, read poses for the detection packet).  Compiled as C11 with
 * -Wall -Wextra -pedantic (the reference's flags, CMakeLists.txt:11) and linked against
 * libsslsim_b200.so by tests/test_abi.py; run on the GPU by tests/test_gpu_parity.py.
 * Prints one line per body so the test can compare with the oracle.
 */
#include <stdio.h>
#include <stdlib.h>
#include "sslsim.h"
#include "sslsim_batch.h"
#include "serialize.h"

int main(int argc, char **argv) {
  int steps = argc > 1 ? atoi(argv[1]) : 60;
  struct World *world = new_world(&FIELD_2015);
  if (!world) return 3;
  for (int i = 0; i < 6; i++) {
    world_add_robot(world, i, TEAM_BLUE);
    world_add_robot(world, i, TEAM_YELLOW);
  }
  for (int s = 0; s < steps; s++) world_step(world);
  printf("frame %u t %.9g field %g x %g limits %g %g\n", world_get_frame_number(world), world_get_timestamp(world),
         (double)world_get_field(world)->field_length, (double)world_get_field(world)->field_width,
         (double)field_limit_x(world_get_field(world)), (double)field_limit_y(world_get_field(world)));
  for (int i = 0; i < world_ball_count(world); i++) {
    struct Vec3 v = ball_get_vec(world_get_ball(world, i));
    printf("ball %.9g %.9g %.9g\n", (double)v.x, (double)v.y, (double)v.z);
  }
  for (int i = 0; i < world_robot_count(world); i++) {
    const struct Robot *r = world_get_robot(world, i);
    struct Pos2 p = robot_get_pos(r);
    printf("robot %d %d %.9g %.9g %.9g\n", get_id(r), (int)get_team(r), (double)p.x, (double)p.y, (double)p.w);
  }
  {
    /* the packet path of the reference loop (src/main.c:83-91) */
    char buffer[10240];
    int n = serialize_world(world, buffer, (int)sizeof buffer);
    int g = serialize_field(world_get_field(world), buffer, (int)sizeof buffer);
    printf("packets %d %d\n", n, g);
  }
  delete_world(world);
  return 0;
}
