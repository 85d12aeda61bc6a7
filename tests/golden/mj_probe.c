/* Golden-vector generator helper (runs only in the build container, needs /root/reference).
 * Links the reference's vendored MuJoCo 3.0.1 binary (mujoco/lib/libmujoco.so.3.0.1, headers mujoco/include) and
 * prints rigid-body quantities of the Hunter MJCF for (qpos, qvel) pairs read from stdin.
 * Build recipe: tests/golden/gen_rbd_mujoco.py (output binary under oracle/_ref/, git-ignored). */
#include <stdio.h>
#include <stdlib.h>
#include <mujoco/mujoco.h>

int main(int argc, char** argv) {
  char err[1000];
  mjModel* m = mj_loadXML(argv[1], NULL, err, 1000);
  if (!m) { fprintf(stderr, "load error: %s\n", err); return 1; }
  mjData* d = mj_makeData(m);
  const char* sites[4] = {"leg_l_f1_site", "leg_r_f1_site", "leg_l_f2_site", "leg_r_f2_site"};
  int nv = m->nv, nq = m->nq;
  printf("%d %d %.17g\n", nq, nv, (double)m->body_subtreemass[1]);
  double* M = (double*)malloc(sizeof(double) * nv * nv);
  double* jp = (double*)malloc(sizeof(double) * 3 * nv);
  while (1) {
    for (int i = 0; i < nq; ++i) if (scanf("%lf", &d->qpos[i]) != 1) return 0;
    for (int i = 0; i < nv; ++i) if (scanf("%lf", &d->qvel[i]) != 1) return 0;
    for (int i = 0; i < nv; ++i) d->qacc[i] = 0;
    mj_forward(m, d);
    mj_subtreeVel(m, d);
    mj_fullM(m, M, d->qM);
    for (int i = 0; i < nv * nv; ++i) printf("%.17g ", M[i]);
    for (int i = 0; i < nv; ++i) printf("%.17g ", d->qfrc_bias[i]);
    for (int s = 0; s < 4; ++s) {
      int id = mj_name2id(m, mjOBJ_SITE, sites[s]);
      for (int i = 0; i < 3; ++i) printf("%.17g ", d->site_xpos[3 * id + i]);
      mj_jacSite(m, d, jp, NULL, id);
      for (int i = 0; i < 3 * nv; ++i) printf("%.17g ", jp[i]);
    }
    for (int i = 0; i < 3; ++i) printf("%.17g ", d->subtree_com[3 + i]);
    for (int i = 0; i < 3; ++i) printf("%.17g ", d->subtree_linvel[3 + i]);
    for (int i = 0; i < 3; ++i) printf("%.17g ", d->subtree_angmom[3 + i]);
    printf("\n");
    fflush(stdout);
  }
  return 0;
}
