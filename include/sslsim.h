/*
 * sslsim.h -- the single-world C API of roboime/ssl-sim, served by the B200
 * batched world-stepper (libsslsim_b200.so).
 *
 * Every declaration below is ABI-identical to the reference declaration whose
 * file:line is cited beside it, so host C/C++ code written against the
 * reference (e.g. reference src/main.c, src/serialize.cpp) links unchanged.
 * The reference spreads these over world.h / robot.h / ball.h / team.h /
 * vector.h / field.h (umbrella: reference src/sslsim.h:11-16); here they live
 * in one header and the six per-topic names are forwarding headers.
 *
 * A `struct World` is a batch of one world on a CUDA device (or one member of
 * a `struct WorldBatch`, see sslsim_batch.h).  There is no CPU implementation
 * behind these calls: without a usable CUDA device `new_world` prints the CUDA
 * error to stderr and returns NULL.
 */
#ifndef SSLSIM_B200_SSLSIM_H
#define SSLSIM_B200_SSLSIM_H
/* keep the reference's per-file guards defined so mixed include orders agree */
#define SSLSIM_H
#define WORLD_H
#define ROBOT_H
#define BALL_H
#define TEAM_H
#define VECTOR_H
#define FIELD_H

#ifdef __cplusplus
extern "C" {
#endif

/* ---- scalar and small POD types: reference src/vector.h:16-36 ------------- */
#ifdef DOUBLE_PRECISON /* spelling as in reference src/vector.h:16; unsupported: the device path is fp32 */
#error "sslsim_b200 computes in fp32; DOUBLE_PRECISON is not supported"
#endif
typedef float Float;                                 /* vector.h:19 */
struct Vec2 { Float x, y; };                         /* vector.h:22-24 */
struct Vec3 { Float x, y, z; };                      /* vector.h:26-28 */
struct Pos2 { Float x, y, w; };                      /* vector.h:30-32; w = yaw [rad] */
struct Pos3 { Float x, y, z, wx, wy, wz; };          /* vector.h:34-36 */

/* ---- teams: reference src/team.h:16 ---------------------------------------- */
enum Team { TEAM_NONE = -1, TEAM_BLUE = 0, TEAM_YELLOW = 1 };

/* ---- field geometry: reference src/field.h:18-46 ---------------------------- */
struct FieldGeometry {
  Float line_width;
  Float field_length;
  Float field_width;
  Float boundary_width;
  Float referee_width;
  Float goal_width;
  Float goal_depth;
  Float goal_wall_width;
  Float center_circle_radius;
  Float defense_radius;
  Float defense_stretch;
  Float free_kick_from_defense_dist;
  Float penalty_spot_from_field_line_dist;
  Float penalty_line_from_spot_dist;
  Float goal_height;
};
/* x/y position of the outer walls (field.h:36-42) */
static inline Float field_limit_x(const struct FieldGeometry *field) {
  return field->field_length / 2 + field->boundary_width + field->referee_width;
}
static inline Float field_limit_y(const struct FieldGeometry *field) {
  return field->field_width / 2 + field->boundary_width + field->referee_width;
}
extern const struct FieldGeometry FIELD_2014_SINGLE; /* field.h:44, values field.c:11-27 */
extern const struct FieldGeometry FIELD_2014_DOUBLE; /* field.h:45, values field.c:29-45 */
extern const struct FieldGeometry FIELD_2015;        /* field.h:46, values field.c:47-63 */
/* Not in the reference: Division-A field for the 11v11 workload. */
extern const struct FieldGeometry FIELD_DIV_A;

struct World;
struct Robot;
struct Ball;
struct btDiscreteDynamicsWorld;
struct btRigidBody;

/* ---- world: reference src/world.h:21-46, bodies src/sslsim.cpp:301-376 ------ */
struct World *new_world(const struct FieldGeometry *field);        /* world.h:21; field is borrowed */
struct World *clone_world(const struct World *world);              /* world.h:22; a real deep copy here */
void delete_world(struct World *world);                            /* world.h:23 */
void world_step(struct World *world);                              /* world.h:25 = step_delta(1/60, 2, 1/120) */
void world_step_delta(struct World *world, Float time_step, int max_substeps,
                      Float fixed_time_step);                      /* world.h:26-27 */
const struct FieldGeometry *world_get_field(const struct World *world);   /* world.h:29 */
int world_ball_count(const struct World *world);                          /* world.h:31 */
const struct Ball *world_get_ball(const struct World *world, int index);  /* world.h:32 */
const struct Ball *world_get_game_ball(const struct World *world);        /* world.h:33 */
struct Ball *world_get_mut_ball(struct World *world, int index);          /* world.h:34 */
struct Ball *world_get_mut_game_ball(struct World *world);                /* world.h:35 */
int world_robot_count(const struct World *world);                         /* world.h:38 */
const struct Robot *world_get_robot(const struct World *world, int index);/* world.h:39 */
struct Robot *world_get_mut_robot(struct World *world, int index);        /* world.h:40 */
void world_add_robot(struct World *world, int id, enum Team team);        /* world.h:41 */
unsigned int world_get_frame_number(const struct World *world);           /* world.h:43 */
double world_get_timestamp(const struct World *world);                    /* world.h:44 */
/* Bullet escape hatch (world.h:46): there is no Bullet here; always NULL. */
struct btDiscreteDynamicsWorld *world_bt_dynamics(struct World *world);

/* ---- ball: reference src/ball.h:18-40, bodies src/sslsim.cpp:378-431 -------- */
struct Vec3 ball_get_vec(const struct Ball *ball);                  /* ball.h:18 */
void ball_set_vec(struct Ball *ball, const struct Vec2 vec);        /* ball.h:19; re-drops from 0.5 m */
/* ball.h:21-26 are TODO stubs upstream (sslsim.cpp:390-405); implemented here:
 * Pos3 = {x,y,z, wx,wy,wz}; pose w* are always 0 (a sphere has no observable
 * attitude), velocity w* is the spin (non-zero only in the ball-spin model). */
struct Pos3 ball_get_pos(const struct Ball *ball);
void ball_set_pos(struct Ball *ball, const struct Pos3 pos);
struct Pos3 ball_get_vel(const struct Ball *ball);
void ball_set_vel(struct Ball *ball, const struct Pos3 vel);
int ball_is_touching_robot(const struct Ball *ball, const struct Robot *robot); /* ball.h:30; during the last step */
struct Robot *ball_last_touching_robot(struct Ball *ball);          /* ball.h:31; NULL if none yet */
Float ball_get_speed2(const struct Ball *ball);                     /* ball.h:34 */
Float ball_get_peak_speed2_from_last_kick(const struct Ball *ball); /* ball.h:37 */
struct btRigidBody *ball_bt_rigid_body(struct Ball *ball);          /* ball.h:40; always NULL */

/* ---- robot: reference src/robot.h:19-33, bodies src/sslsim.cpp:433-471 ------ */
int get_id(const struct Robot *robot);                              /* robot.h:19 */
enum Team get_team(const struct Robot *robot);                      /* robot.h:20 */
struct Pos2 robot_get_pos(const struct Robot *robot);               /* robot.h:22 */
void robot_set_pos(struct Robot *robot, const struct Pos2 pos);     /* robot.h:23; re-drops from 0.5 m */
/* robot.h:25-33 are TODO stubs upstream (sslsim.cpp:452-471); implemented here */
struct Pos2 robot_get_vel(const struct Robot *robot);               /* {vx, vy, yaw rate} */
void robot_set_vel(struct Robot *robot, const struct Pos2 vel);
int robot_is_touching_robot(const struct Robot *robot, const struct Robot *tobor); /* current poses */
Float robot_get_speed2(const struct Robot *robot);

#ifdef __cplusplus
}
#endif
#endif
