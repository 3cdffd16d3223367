// Host-side window builder of the C ABI (include/cpi_b200.h): turns one IMU stream + the update (camera) times into the CSR batch
// layout the preintegration kernels consume.  It replays what the reference does per camera frame:
//   GraphSolver::addmeasurement_imu   (solvers/GraphSolver.cpp:58-69)     readings are queued as they arrive
//   SimulationLoader::execute_publishing (sim/SimulationLoader.cpp:214-290) at equal stamps the IMU reading is delivered first
//   GraphSolver::trytoinitalize       (solvers/GraphSolver.cpp:264, 357)  the first frame that finds >= imuWait readings only
//                                                                          drops all but the newest reading
//   createimufactor_cpi_v1/_v2        (solvers/GraphSolver_IMU.cpp:50-69, 105-124)
//        while (queue.size() > 1 && queue[1].t <= update) { dt = queue[1].t - queue[0].t; if (dt >= 0) feed(queue[0], dt); pop; }
//        dt_f = update - queue[0].t;  if (dt_f > 0) { feed(queue[0], dt_f); queue[0].t = update; }
// No device work: plain C++ compiled into the same library so that C++ callers get it through the same header.
#include <cstdint>

#include "../../include/cpi_b200.h"

namespace cpi { int capi_fail(int code, const char* fmt, ...); }

extern "C" int64_t cpi_cut_windows(int64_t n_imu, const double* t, const double* w, const double* a, int64_t n_updates, const double* update_times,
                                   int64_t imu_wait, int64_t cap_entries, double* samples, int64_t* offsets, int64_t* n_entries_out) {
    if (n_imu < 0 || n_updates < 0 || imu_wait < 0) return cpi::capi_fail(CPI_EINVAL, "negative count");
    if ((n_imu > 0 && (!t || !w || !a)) || (n_updates > 0 && !update_times) || !offsets) return cpi::capi_fail(CPI_EINVAL, "null pointer argument");
    for (int64_t k = 1; k < n_imu; k++)
        if (t[k] < t[k - 1]) return cpi::capi_fail(CPI_EINVAL, "IMU stamps must be non-decreasing (reading %lld)", (long long)k);
    int64_t front = 0;            // index of the reading at the front of the reference's deque
    int64_t arrived = 0;          // readings delivered so far (stamp <= current update time)
    double t_front = n_imu > 0 ? t[0] : 0.0;   // the front's stamp; the partial tail step overwrites it (GraphSolver_IMU.cpp:67)
    bool initialised = imu_wait == 0;
    int64_t nwin = 0, ne = 0;
    offsets[0] = 0;
    auto emit = [&](int64_t i, double dt) {
        if (samples && ne < cap_entries) {
            double* s = samples + ne * CPI_SAMPLE_DOUBLES;
            s[0] = w[3 * i]; s[1] = w[3 * i + 1]; s[2] = w[3 * i + 2]; s[3] = a[3 * i]; s[4] = a[3 * i + 1]; s[5] = a[3 * i + 2]; s[6] = dt;
        }
        ne++;
    };
    for (int64_t u = 0; u < n_updates; u++) {
        const double ut = update_times[u];
        if (u > 0 && ut < update_times[u - 1]) return cpi::capi_fail(CPI_EINVAL, "update times must be non-decreasing (update %lld)", (long long)u);
        while (arrived < n_imu && t[arrived] <= ut) arrived++;
        if (arrived - front < 2) continue;                                   // addmeasurement_uv: "if (imu_times.size() < 2) return"
        if (!initialised) {
            if (arrived - front < imu_wait) continue;
            front = arrived - 1; t_front = t[front]; initialised = true;     // keep the newest reading only
            continue;
        }
        while (arrived - front > 1 && t[front + 1] <= ut) {
            const double dt = t[front + 1] - t_front;
            if (dt >= 0) emit(front, dt);
            front++; t_front = t[front];
        }
        const double dtf = ut - t_front;
        if (dtf > 0) { emit(front, dtf); t_front = ut; }
        offsets[++nwin] = ne;
    }
    if (n_entries_out) *n_entries_out = ne;
    if (samples && ne > cap_entries) return cpi::capi_fail(CPI_ENOMEM, "sample buffer too small: %lld entries needed, capacity %lld", (long long)ne, (long long)cap_entries);
    return nwin;
}
