/*
 * ImuFactorGpu.h -- C++ drop-in IMU factors for the reference tree (rpng/cpi), backed by libcpi_b200.so.
 *
 *     gtsam::ImuFactorCPIv1Gpu   replaces   gtsam::ImuFactorCPIv1   (cpi_compare/src/gtsam/ImuFactorCPIv1.h:55-186)
 *     gtsam::ImuFactorCPIv2Gpu   replaces   gtsam::ImuFactorCPIv2   (cpi_compare/src/gtsam/ImuFactorCPIv2.h:55-205)
 *
 * Same base class (NoiseModelFactor2<JPLNavState, JPLNavState>), same constructor argument lists (ImuFactorCPIv1.h:78-82,
 * ImuFactorCPIv2.h:82-86), same noise model (noiseModel::Gaussian::Covariance(covariance)), same accessors, print / equals /
 * operator<<, and the same
 *     gtsam::Vector evaluateError(const JPLNavState&, const JPLNavState&, boost::optional<Matrix&> H1, boost::optional<Matrix&> H2) const
 * (ImuFactorCPIv1.h:139-140) returning the UNWHITENED residual and Jacobians (GTSAM whitens outside evaluateError).
 *
 * The arithmetic of evaluateError (ImuFactorCPIv1.cpp:37-208, ImuFactorCPIv2.cpp:38-212) runs in the batched CUDA kernel behind
 * cpi_imu_factor_eval_batch_host (include/cpi_b200.h).  Two ways to use it:
 *   (1) per factor, exactly like the reference: every evaluateError() call evaluates this one factor on the device;
 *   (2) graph-level pre-evaluation: cpi_b200::ImuFactorBatch<F>::evaluate(factors, values) evaluates ALL factors of a graph
 *       for the current values in ONE launch and parks each factor's row in the factor; the evaluateError() calls GTSAM then makes
 *       (NoiseModelFactor::linearize, on whatever thread) are served from that row as long as the two states are the ones the
 *       batch was evaluated at (compared bit for bit), and fall back to (1) otherwise.
 * There is no CPU path: errors of the library surface as std::runtime_error.
 *
 * Compiles against real GTSAM in the reference tree; this repository tests it against oracle/gtsam_stub (GTSAM is not installed
 * here), next to the reference's own ImuFactorCPIv1.cpp / ImuFactorCPIv2.cpp in one binary (tests/cpp/test_factor_facade.cpp).
 */
#ifndef CPI_B200_IMUFACTORGPU_H
#define CPI_B200_IMUFACTORGPU_H

#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <gtsam/nonlinear/NonlinearFactor.h>

#include "JPLNavState.h"
#include "cpi_b200.h"

namespace cpi_b200 {

/// One evaluated row: residual and Jacobians of one factor at one pair of states
struct FactorRow {
    double xi[CPI_STATE_DOUBLES], xj[CPI_STATE_DOUBLES];     // the states the row was evaluated at
    double e[15], H1[225], H2[225];                          // column-major, as the C ABI writes them
    bool valid = false;
};

inline void pack_state(const gtsam::JPLNavState& s, double* x) {   // JPLNavState.h:62-66 -> [q(4) bg(3) v(3) ba(3) p(3)]
    for (int k = 0; k < 4; k++) x[k] = s.q()(k);
    for (int k = 0; k < 3; k++) { x[4 + k] = s.bg()(k); x[7 + k] = s.v()(k); x[10 + k] = s.ba()(k); x[13 + k] = s.p()(k); }
}

inline void throw_on(int rc) {
    if (rc != CPI_OK) throw std::runtime_error(std::string("cpi_b200: ") + cpi_last_error());
}

}  // namespace cpi_b200

namespace gtsam {

/// Common part of the two factors: measurement storage in the C-ABI record layout + the evaluation plumbing
class ImuFactorCPIGpuBase : public NoiseModelFactor2<JPLNavState, JPLNavState> {
public:
    typedef Eigen::Matrix<double, 4, 1> JPLQuaternion;
    typedef Eigen::Vector3d Bias3;

protected:
    int model_;
    std::vector<double> rec_;          // CPI_REC_V1_DOUBLES / CPI_REC_V2_DOUBLES: the factor's constructor arguments in record order
    double lin_[CPI_LIN_DOUBLES];      // [bg_lin ba_lin q_K_lin grav]
    mutable std::shared_ptr<cpi_b200::FactorRow> row_;    // parked by ImuFactorBatch::evaluate (graph-level pre-evaluation)

    ImuFactorCPIGpuBase(int model, Key i, Key j, const Eigen::Matrix<double, 15, 15>& covariance)
        : NoiseModelFactor2<JPLNavState, JPLNavState>(noiseModel::Gaussian::Covariance(covariance), i, j), model_(model),
          rec_((size_t)cpi_record_doubles(model), 0.0) {
        std::memset(lin_, 0, sizeof lin_);
        Eigen::Map<Eigen::Matrix<double, 15, 15> >(rec_.data() + CPI_REC_P) = covariance;
        rec_[CPI_REC_R] = rec_[CPI_REC_R + 4] = rec_[CPI_REC_R + 8] = 1.0;     // R_k2tau is not consumed by evaluateError
    }
    void set3(int off, const Eigen::Vector3d& v) { for (int k = 0; k < 3; k++) rec_[off + k] = v(k); }
    void set33(int off, const Eigen::Matrix<double, 3, 3>& m) { Eigen::Map<Eigen::Matrix<double, 3, 3> >(rec_.data() + off) = m; }
    Eigen::Vector3d get3(int off) const { return Eigen::Map<const Eigen::Vector3d>(rec_.data() + off); }

    /// ImuFactorCPIv1.cpp:37 / ImuFactorCPIv2.cpp:38, on the device
    gtsam::Vector evaluate(const JPLNavState& state_i, const JPLNavState& state_j, boost::optional<Matrix&> H1, boost::optional<Matrix&> H2) const {
        double x[2 * CPI_STATE_DOUBLES];
        cpi_b200::pack_state(state_i, x);
        cpi_b200::pack_state(state_j, x + CPI_STATE_DOUBLES);
        std::shared_ptr<cpi_b200::FactorRow> row = row_;          // one atomic-ish copy: const and thread-safe like the reference's method
        gtsam::Vector e(15);
        if (row && row->valid && std::memcmp(row->xi, x, sizeof row->xi) == 0 && std::memcmp(row->xj, x + CPI_STATE_DOUBLES, sizeof row->xj) == 0) {
            e = Eigen::Map<const Eigen::Matrix<double, 15, 1> >(row->e);
            if (H1) *H1 = Eigen::Map<const Eigen::Matrix<double, 15, 15> >(row->H1);
            if (H2) *H2 = Eigen::Map<const Eigen::Matrix<double, 15, 15> >(row->H2);
            return e;
        }
        double eo[15], h1[225], h2[225];
        cpi_b200::throw_on(cpi_imu_factor_eval_batch_host(model_, 1, 2, x, nullptr, nullptr, rec_.data(), lin_, eo, H1 ? h1 : nullptr, H2 ? h2 : nullptr));
        e = Eigen::Map<const Eigen::Matrix<double, 15, 1> >(eo);
        if (H1) *H1 = Eigen::Map<const Eigen::Matrix<double, 15, 15> >(h1);
        if (H2) *H2 = Eigen::Map<const Eigen::Matrix<double, 15, 15> >(h2);
        return e;
    }

public:
    int model() const { return model_; }
    const double* record() const { return rec_.data(); }
    const double* linearization() const { return lin_; }
    void park(const std::shared_ptr<cpi_b200::FactorRow>& r) const { row_ = r; }

    double dt() const { return rec_[CPI_REC_DT]; }
    Vector3 m_alpha() const { return get3(CPI_REC_ALPHA); }
    Vector3 m_beta() const { return get3(CPI_REC_BETA); }
    JPLQuaternion m_q() const { return Eigen::Map<const JPLQuaternion>(rec_.data() + CPI_REC_Q); }
    Bias3 m_balin() const { return Eigen::Map<const Bias3>(lin_ + 3); }
    Bias3 m_bglin() const { return Eigen::Map<const Bias3>(lin_); }
    Bias3 gravity() const { return Eigen::Map<const Bias3>(lin_ + 10); }

    bool equals_measurement(const ImuFactorCPIGpuBase& o, double tol) const {
        if (model_ != o.model_) return false;
        const int nmeas = CPI_REC_P;                      // everything before P_meas (the noise model compares the covariance)
        for (int k = 0; k < nmeas; k++) if (!(std::fabs(rec_[k] - o.rec_[k]) <= tol)) return false;
        for (size_t k = CPI_REC_V1_DOUBLES; k < rec_.size(); k++) if (!(std::fabs(rec_[k] - o.rec_[k]) <= tol)) return false;
        for (int k = 0; k < CPI_LIN_DOUBLES; k++) if (!(std::fabs(lin_[k] - o.lin_[k]) <= tol)) return false;
        return true;
    }
};

/// Drop-in for ImuFactorCPIv1 (ImuFactorCPIv1.h:55)
class ImuFactorCPIv1Gpu : public ImuFactorCPIGpuBase {
public:
    /// ImuFactorCPIv1.h:78-82
    ImuFactorCPIv1Gpu(Key state_i, Key state_j, Eigen::Matrix<double, 15, 15> covariance, double deltatime,
                      Vector3 grav, Vector3 alpha, Vector3 beta, JPLQuaternion q_KtoK1, Bias3 ba_lin, Bias3 bg_lin,
                      Eigen::Matrix<double, 3, 3> J_q, Eigen::Matrix<double, 3, 3> J_beta, Eigen::Matrix<double, 3, 3> J_alpha,
                      Eigen::Matrix<double, 3, 3> H_beta, Eigen::Matrix<double, 3, 3> H_alpha)
        : ImuFactorCPIGpuBase(1, state_i, state_j, covariance) {
        rec_[CPI_REC_DT] = deltatime;
        set3(CPI_REC_ALPHA, alpha); set3(CPI_REC_BETA, beta);
        for (int k = 0; k < 4; k++) rec_[CPI_REC_Q + k] = q_KtoK1(k);
        set33(CPI_REC_JQ, J_q); set33(CPI_REC_JB, J_beta); set33(CPI_REC_JA, J_alpha); set33(CPI_REC_HB, H_beta); set33(CPI_REC_HA, H_alpha);
        for (int k = 0; k < 3; k++) { lin_[k] = bg_lin(k); lin_[3 + k] = ba_lin(k); lin_[10 + k] = grav(k); }
        lin_[9] = 1.0;
    }

    /// ImuFactorCPIv1.h:139-140
    gtsam::Vector evaluateError(const JPLNavState& state_i, const JPLNavState& state_j,
                                boost::optional<Matrix&> H1 = boost::none, boost::optional<Matrix&> H2 = boost::none) const {
        return evaluate(state_i, state_j, H1, H2);
    }

    GTSAM_EXPORT friend std::ostream& operator<<(std::ostream& os, const ImuFactorCPIv1Gpu& f) {      // ImuFactorCPIv1.h:145-154
        os << "dt:[" << f.dt() << "]'" << std::endl;
        os << "alpha:[" << f.m_alpha()(0) << ", " << f.m_alpha()(1) << ", " << f.m_alpha()(2) << "]'" << std::endl;
        os << "beta:[" << f.m_beta()(0) << ", " << f.m_beta()(1) << ", " << f.m_beta()(2) << "]'" << std::endl;
        os << "dq_KtoK1:[" << f.m_q()(0) << ", " << f.m_q()(1) << ", " << f.m_q()(2) << ", " << f.m_q()(3) << "]'" << std::endl;
        os << "ba_lin:[" << f.m_balin()(0) << ", " << f.m_balin()(1) << ", " << f.m_balin()(2) << "]'" << std::endl;
        os << "bg_lin:[" << f.m_bglin()(0) << ", " << f.m_bglin()(1) << ", " << f.m_bglin()(2) << "]'" << std::endl;
        os << "gravity:[" << f.gravity()(0) << ", " << f.gravity()(1) << ", " << f.gravity()(2) << "]'" << std::endl;
        return os;
    }
    void print(const std::string& s, const KeyFormatter& keyFormatter = DefaultKeyFormatter) const {   // ImuFactorCPIv1.h:157-161
        std::cout << s << "ImuFactorCPIv1(" << keyFormatter(this->key1()) << "," << keyFormatter(this->key2()) << ")" << std::endl;
        std::cout << "  measured: " << std::endl << *this << std::endl;
        this->noiseModel_->print("  noise model: ");
    }
    bool equals(const NonlinearFactor& expected, double tol = 1e-9) const {                            // ImuFactorCPIv1.h:164-182
        const ImuFactorCPIv1Gpu* e = dynamic_cast<const ImuFactorCPIv1Gpu*>(&expected);
        if (e == NULL) return false;
        return NoiseModelFactor2<JPLNavState, JPLNavState>::equals(*e, tol) && equals_measurement(*e, tol);
    }
};

/// Drop-in for ImuFactorCPIv2 (ImuFactorCPIv2.h:55)
class ImuFactorCPIv2Gpu : public ImuFactorCPIGpuBase {
public:
    /// ImuFactorCPIv2.h:82-86
    ImuFactorCPIv2Gpu(Key state_i, Key state_j, Eigen::Matrix<double, 15, 15> covariance, double deltatime,
                      Vector3 grav, Vector3 alpha, Vector3 beta, JPLQuaternion q_KtoK1, JPLQuaternion q_K_lin, Bias3 ba_lin, Bias3 bg_lin,
                      Eigen::Matrix<double, 3, 3> J_q, Eigen::Matrix<double, 3, 3> J_beta, Eigen::Matrix<double, 3, 3> J_alpha,
                      Eigen::Matrix<double, 3, 3> H_beta, Eigen::Matrix<double, 3, 3> H_alpha, Eigen::Matrix<double, 3, 3> O_beta,
                      Eigen::Matrix<double, 3, 3> O_alpha)
        : ImuFactorCPIGpuBase(2, state_i, state_j, covariance) {
        rec_[CPI_REC_DT] = deltatime;
        set3(CPI_REC_ALPHA, alpha); set3(CPI_REC_BETA, beta);
        for (int k = 0; k < 4; k++) { rec_[CPI_REC_Q + k] = q_KtoK1(k); lin_[6 + k] = q_K_lin(k); }
        set33(CPI_REC_JQ, J_q); set33(CPI_REC_JB, J_beta); set33(CPI_REC_JA, J_alpha); set33(CPI_REC_HB, H_beta); set33(CPI_REC_HA, H_alpha);
        set33(CPI_REC_OB, O_beta); set33(CPI_REC_OA, O_alpha);
        for (int k = 0; k < 3; k++) { lin_[k] = bg_lin(k); lin_[3 + k] = ba_lin(k); lin_[10 + k] = grav(k); }
    }
    JPLQuaternion m_qklin() const { return Eigen::Map<const JPLQuaternion>(lin_ + 6); }

    /// ImuFactorCPIv2.h:143-144
    gtsam::Vector evaluateError(const JPLNavState& state_i, const JPLNavState& state_j,
                                boost::optional<Matrix&> H1 = boost::none, boost::optional<Matrix&> H2 = boost::none) const {
        return evaluate(state_i, state_j, H1, H2);
    }

    GTSAM_EXPORT friend std::ostream& operator<<(std::ostream& os, const ImuFactorCPIv2Gpu& f) {
        os << "dt:[" << f.dt() << "]'" << std::endl;
        os << "alpha:[" << f.m_alpha()(0) << ", " << f.m_alpha()(1) << ", " << f.m_alpha()(2) << "]'" << std::endl;
        os << "beta:[" << f.m_beta()(0) << ", " << f.m_beta()(1) << ", " << f.m_beta()(2) << "]'" << std::endl;
        os << "dq_KtoK1:[" << f.m_q()(0) << ", " << f.m_q()(1) << ", " << f.m_q()(2) << ", " << f.m_q()(3) << "]'" << std::endl;
        os << "qk_lin:[" << f.m_qklin()(0) << ", " << f.m_qklin()(1) << ", " << f.m_qklin()(2) << ", " << f.m_qklin()(3) << "]'" << std::endl;
        os << "ba_lin:[" << f.m_balin()(0) << ", " << f.m_balin()(1) << ", " << f.m_balin()(2) << "]'" << std::endl;
        os << "bg_lin:[" << f.m_bglin()(0) << ", " << f.m_bglin()(1) << ", " << f.m_bglin()(2) << "]'" << std::endl;
        os << "gravity:[" << f.gravity()(0) << ", " << f.gravity()(1) << ", " << f.gravity()(2) << "]'" << std::endl;
        return os;
    }
    void print(const std::string& s, const KeyFormatter& keyFormatter = DefaultKeyFormatter) const {
        std::cout << s << "ImuFactorCPIv2(" << keyFormatter(this->key1()) << "," << keyFormatter(this->key2()) << ")" << std::endl;
        std::cout << "  measured: " << std::endl << *this << std::endl;
        this->noiseModel_->print("  noise model: ");
    }
    bool equals(const NonlinearFactor& expected, double tol = 1e-9) const {
        const ImuFactorCPIv2Gpu* e = dynamic_cast<const ImuFactorCPIv2Gpu*>(&expected);
        if (e == NULL) return false;
        return NoiseModelFactor2<JPLNavState, JPLNavState>::equals(*e, tol) && equals_measurement(*e, tol);
    }
};

}  // namespace gtsam

namespace cpi_b200 {

/// Graph-level pre-evaluation: all factors of one model, one kernel launch for the current values.
template <class FACTOR> struct ImuFactorBatch {
    /// `values`: state per key (what GTSAM's Values holds for the JPLNavState keys).  After the call every factor serves
    /// evaluateError(values[key1], values[key2]) from its parked row.
    static void evaluate(const std::vector<const FACTOR*>& factors, const std::map<gtsam::Key, gtsam::JPLNavState>& values) {
        if (factors.empty()) return;
        const int model = factors[0]->model();
        const int rd = cpi_record_doubles(model);
        std::map<gtsam::Key, int64_t> slot;
        std::vector<double> X;
        for (typename std::map<gtsam::Key, gtsam::JPLNavState>::const_iterator it = values.begin(); it != values.end(); ++it) {
            slot[it->first] = (int64_t)(X.size() / CPI_STATE_DOUBLES);
            X.resize(X.size() + CPI_STATE_DOUBLES);
            pack_state(it->second, X.data() + X.size() - CPI_STATE_DOUBLES);
        }
        const size_t n = factors.size();
        std::vector<double> rec(n * (size_t)rd), lin(n * CPI_LIN_DOUBLES), e(n * 15), H1(n * 225), H2(n * 225);
        std::vector<int64_t> ii(n), jj(n);
        for (size_t f = 0; f < n; f++) {
            if (factors[f]->model() != model) throw std::runtime_error("cpi_b200: mixed factor models in one batch");
            std::memcpy(rec.data() + f * rd, factors[f]->record(), sizeof(double) * rd);
            std::memcpy(lin.data() + f * CPI_LIN_DOUBLES, factors[f]->linearization(), sizeof(double) * CPI_LIN_DOUBLES);
            std::map<gtsam::Key, int64_t>::const_iterator a = slot.find(factors[f]->key1()), b = slot.find(factors[f]->key2());
            if (a == slot.end() || b == slot.end()) throw std::runtime_error("cpi_b200: factor key missing from the values");
            ii[f] = a->second; jj[f] = b->second;
        }
        throw_on(cpi_imu_factor_eval_batch_host(model, (int64_t)n, (int64_t)(X.size() / CPI_STATE_DOUBLES), X.data(), ii.data(), jj.data(), rec.data(),
                                                lin.data(), e.data(), H1.data(), H2.data()));
        for (size_t f = 0; f < n; f++) {
            std::shared_ptr<FactorRow> r(new FactorRow);
            std::memcpy(r->xi, X.data() + ii[f] * CPI_STATE_DOUBLES, sizeof r->xi);
            std::memcpy(r->xj, X.data() + jj[f] * CPI_STATE_DOUBLES, sizeof r->xj);
            std::memcpy(r->e, e.data() + f * 15, sizeof r->e);
            std::memcpy(r->H1, H1.data() + f * 225, sizeof r->H1);
            std::memcpy(r->H2, H2.data() + f * 225, sizeof r->H2);
            r->valid = true;
            factors[f]->park(r);
        }
    }
};

}  // namespace cpi_b200
#endif
