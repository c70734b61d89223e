/*
 * ref_shim.cpp -- builds oracle/_ref/libmgcfd_ref.so from the reference's OWN sources.
 * TEST INFRASTRUCTURE ONLY (see oracle_api.h).
 *
 * The elemental kernel headers are #included where they lie under /root/reference
 * (-I/root/reference on the command line; no reference source is copied into this
 * repo).  The reference's main() cannot be built (OP2, HDF5 and MPI are absent,
 * SURVEY.md 8c), so the op_par_loop executor and the driver loop nest are the
 * restatement in op2_seq.inc; everything arithmetic below this line is reference code.
 *
 * The globals are the ones euler3d.cpp:47-57 defines; IDIVIDE is NOT defined
 * (inlined_funcs.h:112 is corrupt under it).
 */
#include <stdio.h>
#include <string.h>
#include <sstream>
#include <string>
#include <cmath>
#include <math.h>

#define op_printf printf

#include "const.h"
#include "structures.h"
#include "inlined_funcs.h"
#include "config.h"
#include "utils.h"

/* euler3d.cpp:47-57 */
double smoothing_coefficient = double(0.2f);
double ff_variable[NVAR];
double ff_flux_contribution_momentum_x[NDIM];
double ff_flux_contribution_momentum_y[NDIM];
double ff_flux_contribution_momentum_z[NDIM];
double ff_flux_contribution_density_energy[NDIM];
int mesh_name;
int levels;
int current_level;
#include "global.h"
config conf;

#include "flux.h"
#include "mg.h"
#include "time_stepping_kernels.h"
#include "validation.h"
#include "unstructured_stream.h"
#include "misc.h"
#include "copy_double_kernel.h"

#include "oracle_api.h"

#define EK(name) name
#define ORC_LIB_NAME "reference"
#define G_FF_VARIABLE ff_variable

/* far-field constants: statement-for-statement what euler3d.cpp:157-189 evaluates,
 * calling the reference's compute_flux_contribution (inlined_funcs.h:70-96) */
static void lib_set_farfield(double *out18)
{
    const double aoa = double(PI / 180.0) * double(deg_angle_of_attack);
    ff_variable[VAR_DENSITY] = double(1.4);
    double ff_p = double(1.0);
    double ff_c = sqrt(GAMMA * ff_p / ff_variable[VAR_DENSITY]);
    double ff_speed = double(ff_mach) * ff_c;
    double3 ff_v;
    ff_v.x = ff_speed * double(cos((double)aoa));
    ff_v.y = ff_speed * double(sin((double)aoa));
    ff_v.z = 0.0;
    ff_variable[VAR_MOMENTUM + 0] = ff_variable[VAR_DENSITY] * ff_v.x;
    ff_variable[VAR_MOMENTUM + 1] = ff_variable[VAR_DENSITY] * ff_v.y;
    ff_variable[VAR_MOMENTUM + 2] = ff_variable[VAR_DENSITY] * ff_v.z;
    ff_variable[VAR_DENSITY_ENERGY] =
        ff_variable[VAR_DENSITY] * (double(0.5) * (ff_speed * ff_speed)) + (ff_p / double(GAMMA - 1.0));
    double3 ff_m;
    ff_m.x = ff_variable[VAR_MOMENTUM + 0];
    ff_m.y = ff_variable[VAR_MOMENTUM + 1];
    ff_m.z = ff_variable[VAR_MOMENTUM + 2];
    compute_flux_contribution(ff_variable[VAR_DENSITY], ff_m, ff_variable[VAR_DENSITY_ENERGY], ff_p, ff_v,
                              ff_flux_contribution_momentum_x, ff_flux_contribution_momentum_y,
                              ff_flux_contribution_momentum_z, ff_flux_contribution_density_energy);
    if (out18) {
        out18[0] = smoothing_coefficient;
        for (int i = 0; i < 5; i++) out18[1 + i] = ff_variable[i];
        for (int i = 0; i < 3; i++) {
            out18[6 + i] = ff_flux_contribution_momentum_x[i];
            out18[9 + i] = ff_flux_contribution_momentum_y[i];
            out18[12 + i] = ff_flux_contribution_momentum_z[i];
            out18[15 + i] = ff_flux_contribution_density_energy[i];
        }
    }
}

extern "C" {
#include "op2_seq.inc"
}
