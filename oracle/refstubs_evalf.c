/*
 * refstubs_evalf.c -- TEST INFRASTRUCTURE ONLY (oracle/_ref build of libmoldyref_evalf.so).
 *
 * accel.c (compiled *in place* from /root/reference/src, never copied) defines eval_forces() -- the oracle of
 * SURVEY 8f rank 1 -- next to do_step() and rescale(), which import the averages, dump and thermalise
 * subsystems.  Those are outside the hot-path contract and are never reached through eval_forces(); the
 * symbols below only satisfy the dynamic linker and abort when called.
 *
 * Nothing under moldy_b200/ may link or load this file.
 */
#include <stdio.h>
#include <stdlib.h>

static void unreachable(const char *name)
{
   fprintf(stderr, "oracle: %s() is outside the eval_forces() contract\n", name);
   abort();
}
double value(void)      { unreachable("value");      return 0.0; }
double roll_av(void)    { unreachable("roll_av");    return 0.0; }
void   dump(void)       { unreachable("dump"); }
void   thermalise(void) { unreachable("thermalise"); }
