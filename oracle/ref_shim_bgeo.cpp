// TEST INFRASTRUCTURE ONLY -- never linked into or called from the product path.
//
// The reference's IO_CUDA writer compiled IN PLACE: save_convex_cells_houdini + get_one_convex_cell_faces_const
// (src/IO/IO_CUDA/io_cuda.cxx:21-187) and IO::GeometryWriter (src/IO/IO_CUDA/io_utils.cpp), on top of
// ConvexCellHost (src/rpd3d_base/voronoi_defs.cxx).  geogram / nlohmann_json / matbase / inputs headers are
// replaced by the minimal stand-ins in oracle/stubs (the writer uses a 3-vector, two empty structs and three
// helpers from them).  Records come in the ConvexCellTransfer layout and are expanded with the copy_cc rule
// (src/rpd3d/voronoi.cu:433-449) like oracle/ref_shim_host.cpp does.
#include <signal.h>
#include <sys/wait.h>
#include <unistd.h>

#include <cstring>
#include <stdexcept>
#include <string>

#include "voronoi_defs.cxx"  // -I/root/reference/src/rpd3d_base
#include "io_utils.cpp"      // -I/root/reference/src/IO/IO_CUDA
#include "io_cuda.cxx"

#include "oracle.h"

static void expand(const orc_record& r, ConvexCellHost& h, int id) {
  h.is_active = true;
  h.status = (Status)r.status;
  h.thread_id = r.thread_id;
  h.voro_id = r.voro_id;
  h.tet_id = r.tet_id;
  h.euler = r.euler;
  h.weight = r.weight;
  h.nb_v = r.nb_v;
  h.nb_p = r.nb_p;
  h.nb_e = r.nb_e;
  for (int i = 0; i < r.nb_v; i++)
    h.ver_data_trans[i] = cmake_uchar4(r.ver[i][0], r.ver[i][1], r.ver[i][2], r.ver[i][3]);
  for (int i = 0; i < r.nb_p; i++) {
    h.clip_data_trans[i] = cmake_float5(r.clip[i].x, r.clip[i].y, r.clip[i].z, r.clip[i].w, r.clip[i].h);
    h.clip_id2_data_trans[i] = cmake_int2(r.id2[i][0], r.id2[i][1]);
  }
  for (int i = 0; i < r.nb_e; i++) h.edge_data[i] = cmake_uchar3(r.edge[i][0], r.edge[i][1], r.edge[i][2]);
  h.id = id;
}

extern "C" {

// Runs the reference's save_convex_cells_houdini with the working directory set to `work_dir`: the file appears at
// <work_dir>/../out/<name>/rpd/rpd_<name>_ref.bgeo (io_cuda.cxx:158-161 with the stand-in time stamp).
int ref_bgeo_write(const orc_record* recs, long n, int max_sf_fid, int is_boundary_only, const char* work_dir,
                   const char* name) {
  std::vector<ConvexCellHost> cells((size_t)n);
  for (long i = 0; i < n; i++) {
    expand(recs[i], cells[(size_t)i], (int)i);
    cells[(size_t)i].reload_active();  // update_power_cells does this before any writer runs (rpd_update.cxx:625-626)
  }
  // save_convex_cells_houdini is declared bool and falls off its end without a return (io_cuda.cxx:186-187): g++
  // plants a trap there, so the call dies AFTER the file has been written and closed (the stream is a local of
  // GeometryWriter::OutputGeometry).  It therefore runs in a forked child; the parent only waits for it.
  fflush(stdout);
  fflush(stderr);
  const pid_t pid = fork();
  if (pid < 0) return -1;
  if (pid == 0) {
    // the expected trap must not run the host interpreter's fault handlers (pytest's faulthandler prints a dump)
    signal(SIGSEGV, SIG_DFL);
    signal(SIGILL, SIG_DFL);
    signal(SIGABRT, SIG_DFL);
    if (chdir(work_dir) != 0) _exit(2);
    try {
      Parameter params;
      std::vector<MedialSphere> spheres;
      save_convex_cells_houdini(params, spheres, cells, name, max_sf_fid, is_boundary_only != 0, false);
    } catch (const std::exception& e) {
      fprintf(stderr, "ref_bgeo_write: %s\n", e.what());
      _exit(3);
    }
    _exit(0);
  }
  int status = 0;
  waitpid(pid, &status, 0);
  if (WIFEXITED(status) && (WEXITSTATUS(status) == 2 || WEXITSTATUS(status) == 3)) return -WEXITSTATUS(status);
  return 0;  // exit 0 or the trap after the write: the caller checks the file
}

}  // extern "C"
