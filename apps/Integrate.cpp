// Integrate -- drop-in for the reference's Integrate executable (Integrate/Integrate.cpp, IntegrateApp.cpp):
// same command line, same .log / camera / .ctr inputs, same world.pcd output; the TSDF volume lives on a B200
// behind libelastic_b200 (include/elastic_b200.h).
//
// Host logic restated from CIntegrateApp::{Init,StartMainLoop,Execute} (IntegrateApp.cpp:43-141,190-226); the
// numeric work (ScaleDepth, Integrate, Reproject, SaveWorld) is the library's.  Differences, all at the edges:
//   * "-oni <file>" names the raw depth container of apps/er_io.h (OpenNI is not available); "-dev" cannot open;
//   * frames are pulled synchronously, so there is no 10 x 300 ms end-of-stream timeout;
//   * extra flags: --device N, --devices A,B,.. (one volume partitioned by blocks over several GPUs: device k owns
//     the blocks with (xi+yi+zi) % n == k, every device sees every frame, the extracted voxels are concatenated --
//     no halo is needed, a voxel update reads only depth + pose), --max_blocks N, --fusion N (frames per fused
//     launch), --bench_json FILE, and
//     --voxel_size M / --tsdf_trunc M / --block_res 64: the reference's constructor constants
//     (TSDFVolume.cpp:10-11,57), defaulting to the reference values 3/512, 0.03, 64.
#include <chrono>
#include <cstdio>
#include <iostream>
#include <string>
#include <thread>
#include <vector>

#include "elastic_b200.h"
#include "er_io.h"

using namespace er_io;

static int print_help() {
  std::cout << "\nApplication parameters:" << std::endl;
  std::cout << "    --help, -h                      : print this message" << std::endl;
  std::cout << "    --ref_traj <log_file>           : use a reference trajectory file" << std::endl;
  std::cout << "    --pose_traj <log_file>          : use a pose trajectory file to create a reference trajectory" << std::endl;
  std::cout << "    --seg_traj <log_file>           : trajectory within each fragment - must have" << std::endl;
  std::cout << "    --ctr <ctr_file>                : enables distortion, must specify the following parameters" << std::endl;
  std::cout << "    --num <number>                  : number of pieces, important parameter" << std::endl;
  std::cout << "    --resolution <resolution>       : default - 8" << std::endl;
  std::cout << "    --length <length>               : default - 3.0" << std::endl;
  std::cout << "    --interval <interval>           : default - 50" << std::endl;
  std::cout << "    --camera <param_file>           : load camera parameters" << std::endl;
  std::cout << "    --save_to <pcd_file>            : output file, default - world.pcd" << std::endl;
  std::cout << "    --start_from <frame_id>         : frames before frame_id will be skipped" << std::endl;
  std::cout << "    --end_at <frame_id>             : frames after frame_id will be skipped" << std::endl;
  std::cout << "Valid depth data sources:" << std::endl;
  std::cout << "    -dev <device> (default), -oni <oni_file>" << std::endl;
  std::cout << "B200 build only:" << std::endl;
  std::cout << "    --device <n> --max_blocks <n> --fusion <1..32> --bench_json <file>" << std::endl;
  std::cout << "    --save_volume <file> / --load_volume <file[,file...]> : raw volume checkpoint after / before the integration" << std::endl;
  std::cout << "    --devices <a,b,...>             : partition the volume's blocks over several GPUs" << std::endl;
  std::cout << "    --voxel_size <m> (3/512) --tsdf_trunc <m> (0.03) --block_res <n> (64, fixed)" << std::endl;
  return 0;
}

struct App {
  // defaults: CIntegrateApp::CIntegrateApp (IntegrateApp.cpp:5-26)
  std::string traj_filename, pose_filename, seg_filename, camera_filename, ctr_filename, pcd_filename = "world.pcd";
  int ctr_num = 0, ctr_resolution = 8, ctr_interval = 50, start_from = -1, end_at = 100000000;
  double ctr_length = 3.0;
  int device = 0, max_blocks = 4096, fusion = 32, block_res = 64;
  std::vector<int> devices;                                // empty: {device}
  double voxel_size = 0.0, tsdf_trunc = 0.0;               // 0 = the reference constants
  std::string bench_json;

  std::string save_volume, load_volume;   // raw volume checkpoint written after / read before the integration
  CameraParam camera;
  Trajectory traj, seg_traj, pose_traj;
  std::vector<float> grids;
  std::vector<er_tsdf*> vols;                              // one per device; vols[k] owns the blocks of partition k
  int cols = 640, rows = 480;
  int frame_id = 0;
  bool exit_ = false;
  long processed = 0;
  double read_ms = 0;

  bool fail(const char* what) {
    std::cerr << what << ": " << er_last_error() << std::endl;
    return false;
  }

  // CIntegrateApp::Init (IntegrateApp.cpp:43-79)
  bool init() {
    if (file_exists(camera_filename)) camera.load(camera_filename);
    if (ctr_num > 0 && file_exists(ctr_filename) && file_exists(seg_filename)) {
      load_ctr(ctr_filename, ctr_num, ctr_resolution, grids);
    } else {
      ctr_num = 0;
    }
    if (file_exists(traj_filename)) traj.load(traj_filename);
    if (file_exists(seg_filename)) {
      seg_traj.load(seg_filename);
      if (file_exists(pose_filename)) {
        pose_traj.load(pose_filename);
        traj.data.clear();
        for (int i = 0; i < (int)pose_traj.data.size(); i++) {
          for (int j = 0; j < ctr_interval; j++) {
            const int idx = i * ctr_interval + j;
            if (idx >= (int)seg_traj.data.size()) {      // the reference reads out of bounds here (:73)
              std::cerr << "seg_traj has fewer than num*interval entries" << std::endl;
              return false;
            }
            FramedTransformation ft;
            ft.id1 = idx; ft.id2 = idx; ft.frame = idx + 1;
            er_mat4_mul(pose_traj.data[i].T, seg_traj.data[idx].T, ft.T);
            traj.data.push_back(ft);
          }
        }
        std::fprintf(stderr, "Trajectory created from pose and segment trajectories.\n");
      }
    }
    if (devices.empty()) devices.push_back(device);
    const int ndev = (int)devices.size();
    for (int k = 0; k < ndev; k++) {
      er_tsdf* vol = er_tsdf_create(devices[k], cols, rows, voxel_size, block_res, tsdf_trunc, (size_t)max_blocks);
      if (!vol) return fail("er_tsdf_create");
      vols.push_back(vol);
      if (er_tsdf_set_camera(vol, camera.fx, camera.fy, camera.cx, camera.cy, camera.integration_trunc)) return fail("set_camera");
      if (er_tsdf_set_fusion(vol, fusion)) return fail("set_fusion");
      if (ctr_num > 0 && er_tsdf_set_control_grids(vol, grids.data(), ctr_num, ctr_resolution, (float)ctr_length)) return fail("set_control_grids");
      if (ndev > 1 && er_tsdf_set_partition(vol, k, ndev)) return fail("set_partition");
      // --load_volume a[,b,...]: resume from checkpoints; a partitioned volume keeps only the blocks it owns, so the
      // per-device files of an earlier --devices run can be given to any number of devices
      for (size_t pos = 0; pos < load_volume.size();) {
        size_t end = load_volume.find(',', pos);
        if (end == std::string::npos) end = load_volume.size();
        if (end > pos && er_tsdf_load(vol, load_volume.substr(pos, end - pos).c_str())) return fail("er_tsdf_load");
        pos = end + 1;
      }
    }
    return true;
  }

  // --save_volume F: device k of n > 1 writes F.part<k>
  bool save_checkpoint() {
    if (save_volume.empty()) return true;
    for (size_t k = 0; k < vols.size(); k++) {
      const std::string fn = vols.size() == 1 ? save_volume : save_volume + ".part" + std::to_string(k);
      if (er_tsdf_save(vols[k], fn.c_str())) return fail("er_tsdf_save");
    }
    return true;
  }

  // The gating of CIntegrateApp::Execute (IntegrateApp.cpp:190-226) for FrameID `fid`:
  // returns true if the frame is to be integrated; may set exit_.
  bool gate(int fid) {
    if (fid - 1 < 0 || fid - 1 >= (int)traj.data.size()) { exit_ = true; return false; }   // reference: out-of-bounds read
    if (traj.data[fid - 1].frame == -1) return false;
    if (fid >= (int)traj.data.size()) { exit_ = true; return false; }
    if (fid % 100 == 0) std::fprintf(stderr, "Frames processed : %d / %d\n", fid, (int)traj.data.size());
    if (fid < start_from || fid > end_at) {
      if (fid > end_at) { std::fprintf(stderr, "Reaching the specified end point.\n"); exit_ = true; }
      return false;
    }
    if (ctr_num > 0 && fid > ctr_interval * ctr_num) { exit_ = true; return false; }          // Reproject(), :230-233
    return true;
  }

  bool run(DepthStream& src) {
    const size_t npix = (size_t)cols * rows;
    const int chunk_frames = 32;
    // two page-locked chunk buffers (er_host_alloc): the stream is read straight into them and handed to the library
    // without another copy; while the devices still copy chunk n out of one buffer the next read fills the other
    uint16_t* bufs[2] = {(uint16_t*)er_host_alloc(sizeof(uint16_t) * npix * chunk_frames),
                         (uint16_t*)er_host_alloc(sizeof(uint16_t) * npix * chunk_frames)};
    struct Free { uint16_t** b; ~Free() { er_host_free(b[0]); er_host_free(b[1]); } } free_bufs{bufs};
    if (!bufs[0] || !bufs[1]) return fail("er_host_alloc");
    std::vector<double> poses(16 * chunk_frames), segs(16 * chunk_frames);
    std::vector<int> chunks(chunk_frames);
    // the next chunk is read by a helper thread while this one is gated, copied and enqueued
    int which = 0;
    int got_next = 0;
    double rd_ms = 0;
    auto read_into = [&](uint16_t* b) {
      const auto tr0 = std::chrono::steady_clock::now();
      got_next = src.read(b, chunk_frames);
      rd_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tr0).count();
    };
    read_into(bufs[0]);
    while (!exit_) {
      uint16_t* buf = bufs[which];
      const int got = got_next;
      if (got == 0) break;                                    // end of stream
      which ^= 1;
      std::thread reader(read_into, bufs[which]);
      struct Join { std::thread& t; ~Join() { if (t.joinable()) t.join(); } } join_reader{reader};
      int nb = 0;
      for (int i = 0; i < got && !exit_; i++) {
        frame_id++;                                           // FrameID is 1-based (IntegrateApp.cpp:185)
        if (!gate(frame_id)) continue;
        const double* T = traj.data[frame_id - 1].T;
        if (nb != i) std::memmove(buf + npix * nb, buf + npix * i, sizeof(uint16_t) * npix);   // close the gap of skipped frames
        std::memcpy(poses.data() + 16 * nb, T, sizeof(double) * 16);
        if (ctr_num > 0) {                                    // Reproject() inputs of this frame (IntegrateApp.cpp:241-242)
          std::memcpy(segs.data() + 16 * nb, seg_traj.data[frame_id - 1].T, sizeof(double) * 16);
          chunks[nb] = (frame_id - 1) / ctr_interval;
        }
        nb++;
        processed++;
      }
      if (nb > 0) {
        // every device integrates the same frames into its own share of the blocks (one host thread per device); a pool
        // or work-list overflow of an earlier chunk comes back from this call (ER_ERR_CAPACITY) and ends the run
        const int ndev = (int)vols.size();
        std::vector<std::string> errs((size_t)ndev);
#pragma omp parallel for num_threads(ndev) schedule(static, 1)
        for (int k = 0; k < ndev; k++) {
          int rc;
          if (ctr_num > 0)
            rc = er_tsdf_integrate_ctr_batch(vols[k], buf, poses.data(), traj.data[0].T, seg_traj.data[0].T, segs.data(),
                                             chunks.data(), nb);
          else
            rc = er_tsdf_integrate_batch(vols[k], buf, poses.data(), nb);
          if (rc) errs[(size_t)k] = er_last_error();       // the error text is per host thread
        }
        for (int k = 0; k < ndev; k++)
          if (!errs[(size_t)k].empty()) {
            std::cerr << (ctr_num > 0 ? "er_tsdf_integrate_ctr_batch: " : "er_tsdf_integrate_batch: ") << errs[(size_t)k] << std::endl;
            return false;
          }
      }
    }
    read_ms = rd_ms;
    for (er_tsdf* vol : vols)
      if (er_tsdf_sync(vol)) return fail("er_tsdf_sync");
    return true;
  }

  // TSDFVolume::SaveWorld (TSDFVolume.cpp:104-132)
  bool save_world() {
    std::vector<float> pts;
    size_t n = 0;
    for (er_tsdf* vol : vols) {                            // the partitions are disjoint: concatenate
      size_t nk = 0, mk = 0;
      if (er_tsdf_extract(vol, nullptr, 0, &nk)) return fail("er_tsdf_extract");
      pts.resize(4 * (n + nk) + 4);
      if (nk && er_tsdf_extract(vol, pts.data() + 4 * n, nk, &mk)) return fail("er_tsdf_extract");
      n += nk;
    }
    if (!save_pcd_xyzi(pcd_filename, pts.data(), n)) { std::cerr << "cannot write " << pcd_filename << std::endl; return false; }
    std::printf("%d voxel points have been written.\n", (int)n);
    return true;
  }
};

int main(int argc, char* argv[]) {
  if (argc == 1 || find_switch(argc, argv, "--help") || find_switch(argc, argv, "-h")) return print_help();

  std::string openni_device, oni_file;
  DepthStream src;
  if (parse_argument(argc, argv, "-dev", openni_device) > 0) {
    return std::cout << "Can't open depth source" << std::endl, -1;       // no OpenNI in this build
  } else if (parse_argument(argc, argv, "-oni", oni_file) > 0) {
    if (!src.open(oni_file)) return std::cout << "Can't open depth source" << std::endl, -1;
  } else {
    return std::cout << "Can't open depth source" << std::endl, -1;
  }

  App app;
  app.cols = src.w; app.rows = src.h;
  parse_argument(argc, argv, "--ref_traj", app.traj_filename);
  parse_argument(argc, argv, "--pose_traj", app.pose_filename);
  parse_argument(argc, argv, "--seg_traj", app.seg_filename);
  parse_argument(argc, argv, "--camera", app.camera_filename);
  parse_argument(argc, argv, "--save_to", app.pcd_filename);
  parse_argument(argc, argv, "--start_from", app.start_from);
  parse_argument(argc, argv, "--end_at", app.end_at);
  parse_argument(argc, argv, "--ctr", app.ctr_filename);
  parse_argument(argc, argv, "--num", app.ctr_num);
  parse_argument(argc, argv, "--resolution", app.ctr_resolution);
  parse_argument(argc, argv, "--length", app.ctr_length);
  parse_argument(argc, argv, "--interval", app.ctr_interval);
  parse_argument(argc, argv, "--device", app.device);
  std::string devices;
  if (parse_argument(argc, argv, "--devices", devices) > 0) {
    for (size_t pos = 0; pos < devices.size();) {
      size_t end = devices.find(',', pos);
      if (end == std::string::npos) end = devices.size();
      if (end > pos) app.devices.push_back(std::atoi(devices.substr(pos, end - pos).c_str()));
      pos = end + 1;
    }
  }
  parse_argument(argc, argv, "--max_blocks", app.max_blocks);
  parse_argument(argc, argv, "--fusion", app.fusion);
  parse_argument(argc, argv, "--bench_json", app.bench_json);
  parse_argument(argc, argv, "--voxel_size", app.voxel_size);
  parse_argument(argc, argv, "--tsdf_trunc", app.tsdf_trunc);
  parse_argument(argc, argv, "--block_res", app.block_res);
  parse_argument(argc, argv, "--save_volume", app.save_volume);
  parse_argument(argc, argv, "--load_volume", app.load_volume);

  std::cout << "Registration mode: On" << std::endl;
  const auto t0 = std::chrono::steady_clock::now();
  bool ok = app.init();
  const auto t1 = std::chrono::steady_clock::now();
  if (ok) ok = app.run(src);
  const auto t2 = std::chrono::steady_clock::now();
  if (ok) ok = app.save_checkpoint();
  if (ok) ok = app.save_world();
  std::cout << "Total " << app.frame_id << " frames processed." << std::endl;
  const auto t3 = std::chrono::steady_clock::now();
  auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
  std::cerr << "Integrate All took " << ms(t1, t3) << "ms." << std::endl;
  if (!app.bench_json.empty()) {
    if (FILE* f = std::fopen(app.bench_json.c_str(), "w")) {
      std::fprintf(f, "{\"frames_integrated\": %ld, \"init_ms\": %.3f, \"integrate_ms\": %.3f, \"read_ms\": %.3f, \"save_ms\": %.3f, "
                      "\"devices\": %d}\n",
                   app.processed, ms(t0, t1), ms(t1, t2), app.read_ms, ms(t2, t3), (int)app.vols.size());
      std::fclose(f);
    }
  }
  for (er_tsdf* vol : app.vols) er_tsdf_destroy(vol);
  return 0;        // like the reference, failures are reported on the console, not through the exit code
}
