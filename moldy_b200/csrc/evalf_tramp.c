/* evalf_tramp.c -- one strong definition of eval_forces() for a host program that is linked STATICALLY with its own
 * accel.o: compile accel.c with -fPIC (so that do_step's call goes through the symbol), weaken its definition with
 * `objcopy --weaken-symbol=eval_forces accel.o`, and add this object; the linker then binds every call to the version
 * below, which forwards to the library (oracle/Makefile target moldy_gpu_evalf). */
#include "moldy_b200.h"

void eval_forces(system_mp sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, double *pe, real *dip_mom,
                 mat_mt stress, vec_mp *force, vec_mp *torque)
{
   mdb_eval_forces_moldy(sys, species, site_info, potpar, pe, dip_mom, stress, force, torque);
}
