/*
 * ref_field_shim.c -- test infrastructure only.  Compiled TOGETHER WITH the reference's own src/field.c (where it
 * lies under /root/reference; nothing of it is copied) into oracle/_ref/libfield_ref.so by oracle/Makefile.  It
 * exports the reference's `inline` field_limit_x / field_limit_y (src/field.h:36-42), which have no symbol of
 * their own, next to the reference's FIELD_2014_SINGLE / FIELD_2014_DOUBLE / FIELD_2015 (src/field.c:11-63), so
 * that tests/test_reference_field.py can pin the part of the hot path's scene the reference can still build
 *  against the real thing.  The step itself needs Bullet and cannot be built.
 */
#include "field.h" /* the reference's, via -I/root/reference/src */

float ref_field_limit_x(const struct FieldGeometry *f) { return field_limit_x(f); }
float ref_field_limit_y(const struct FieldGeometry *f) { return field_limit_y(f); }
int ref_field_geometry_size(void) { return (int)sizeof(struct FieldGeometry); }
