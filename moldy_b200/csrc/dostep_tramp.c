/* dostep_tramp.c -- link-time trampoline for a host program that takes do_step() (src/accel.c:626-827) from
 * libmoldy_b200.so: NVE dynamics run on the device (mdb_do_step_moldy); for thermostatted / constant-stress runs the
 * program's own do_step -- renamed do_step_host with `objcopy --redefine-sym do_step=do_step_host accel.o` -- still
 * runs on the host and reaches the GPU through eval_forces (evalf_tramp.c).  INTEGRATION.md section 6. */
#include "moldy_b200.h"

extern contr_mt control;
void do_step_host(system_mt *, spec_mt *, site_mt *, pot_mt *, vec_mt (*)[2], double *, real *, mat_mt, void *, int, int);

void do_step(system_mt *sys, spec_mt *species, site_mt *site_info, pot_mt *potpar, vec_mt (*meansq_f_t)[2], double *pe,
             real *dip_mom, mat_mt stress_vir, void *restart_header, int backup_restart, int init_H_0)
{
   if (control.const_temp || control.const_pressure)
      do_step_host(sys, species, site_info, potpar, meansq_f_t, pe, dip_mom, stress_vir, restart_header, backup_restart, init_H_0);
   else
      mdb_do_step_moldy(sys, species, site_info, potpar, meansq_f_t, pe, dip_mom, stress_vir, restart_header, backup_restart,
                        init_H_0);
}
