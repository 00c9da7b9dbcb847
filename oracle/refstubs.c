/*
 * refstubs.c -- TEST INFRASTRUCTURE ONLY (oracle/_ref build).
 *
 * The reference's force.c / kernel.c / ewald.c (compiled *in place* from
 * /root/reference/src by oracle/Makefile, never copied into this repo) import a
 * handful of symbols from the rest of Moldy.  This file supplies them so the
 * three hot-path files can live in a stand-alone shared library that the
 * parity tests load through ctypes:
 *
 *   control, ithread, nthreads   globals owned by main.c        (src/main.c:83-84)
 *   message(), note()            severity-tagged printing       (src/output.c:131,175)
 *   new_line(), put_line()       output helpers rdf.c imports   (src/output.c)
 *
 * rdf.c itself (init_rdf, rdf_accum, rdf_ptr) is compiled in place with the three hot files, so the RDF pass
 * of force_calc bins into the reference's own float histograms.
 *
 * note()/message() additionally record their text in a ring buffer so tests
 * can read the start-up notes (they are the only goldens the reference's own
 * example outputs pin: subcell count, neighbour-cell count, self-energy,
 * k-vector count -- SURVEY.md section 8c).
 *
 * Nothing under moldy_b200/ may link or load this file.
 */
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "defs.h"
#include "structs.h"

contr_mt control;
int ithread = 0, nthreads = 1;

#define LOGCAP (1 << 16)
static char logbuf[LOGCAP];
static size_t loglen = 0;
static int quiet = 1;
static long n_warn = 0, n_err = 0;

static void log_append(const char *tag, const char *fmt, va_list ap)
{
   char line[1024];
   int n = snprintf(line, sizeof line, "%s", tag);
   n += vsnprintf(line + n, sizeof line - (size_t)n - 2, fmt, ap);
   if (n > (int)sizeof line - 2)
      n = (int)sizeof line - 2;
   line[n++] = '\n';
   line[n] = 0;
   if (!quiet)
      fputs(line, stdout);
   if (loglen + (size_t)n < LOGCAP) {
      memcpy(logbuf + loglen, line, (size_t)n + 1);
      loglen += (size_t)n;
   }
}

void note(char *text, ...)
{
   va_list ap;
   if (ithread > 0)
      return;
   va_start(ap, text);
   log_append(" *I* ", text, ap);
   va_end(ap);
}

void message(int *nerrs, ...)
{
   static const char *tag[] = {" *I* ", " *W* ", " *E* ", " *F* "};
   va_list ap;
   char *buff, *fmt;
   int sev;
   va_start(ap, nerrs);
   buff = va_arg(ap, char *);
   sev = va_arg(ap, int);
   fmt = va_arg(ap, char *);
   (void)buff;
   if (abs(sev) == 1) n_warn++;
   if (abs(sev) >= 2) n_err++;
   if (ithread == 0 || abs(sev) == 3)
      log_append(tag[abs(sev) & 3], fmt, ap);
   va_end(ap);
   if (sev >= 2 && nerrs != 0)
      (*nerrs)++;
   if (abs(sev) == 3) {
      fputs(logbuf, stderr);
      fflush(stderr);
      exit(3);
   }
}

void new_line(void) { if (!quiet) putchar('\n'); }
void put_line(int c) { (void)c; }

/* ---- accessors used by the Python harness ---- */
contr_mt *mdref_control(void) { return &control; }
void mdref_set_thread(int it, int nt) { ithread = it; nthreads = nt; }
const char *mdref_log(void) { return logbuf; }
void mdref_log_clear(void) { loglen = 0; logbuf[0] = 0; }
void mdref_set_quiet(int q) { quiet = q; }
long mdref_warnings(void) { return n_warn; }
long mdref_errors(void) { return n_err; }
size_t mdref_sizeof_control(void) { return sizeof(contr_mt); }
size_t mdref_sizeof_system(void) { return sizeof(system_mt); }
size_t mdref_sizeof_spec(void) { return sizeof(spec_mt); }
size_t mdref_sizeof_pot(void) { return sizeof(pot_mt); }
