/* oracle/shim/example_missing_globals.c -- TEST INFRASTRUCTURE ONLY.
 *
 * examples/CT_sc_ductile_nonlocal.c of the reference predates three globals that its current src/ tree
 * declares in include/lpm.h:78 and touches in initialization.c (initMatrices, createCuboid) and
 * data_handler.c (writeBondforce): the example does not define them, so it does not link against the
 * reference's own sources as shipped (SURVEY section 8, config C5src "+3 missing global defs").  The default
 * driver defines them at src/lpmc_project.c:42; this file supplies the same three definitions so that the
 * UNMODIFIED example can be built -- once all-CPU, once against the GPU drop-in library. */
double **bond_stretch, **bond_vector, **bond_force;
