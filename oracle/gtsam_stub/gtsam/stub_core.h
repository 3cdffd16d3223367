// TEST INFRASTRUCTURE ONLY.  Minimal stand-in for the handful of GTSAM / Boost names that the reference's
// ImuFactorCPIv1.cpp, ImuFactorCPIv2.cpp and JPLNavState.cpp touch, so that those translation units can be
// compiled UNMODIFIED, in place from /root/reference, into oracle/_ref/libcpi_ref.so (GTSAM and Boost are not
// installed in this image).  Nothing here does arithmetic: evaluateError() itself only uses
// Eigen and utils/quat_ops.h.  The noise model is a no-op holder: whitening happens outside evaluateError in
// real GTSAM and is out of parity scope ("parity unpinned").
#pragma once
#include <Eigen/Dense>
#include <cstdint>
#include <cmath>
#include <functional>
#include <iostream>
#include <memory>
#include <string>

#define GTSAM_EXPORT

namespace boost {
struct none_t {};
static const none_t none = none_t();
// just enough of boost::optional<T&> for "if (H1) *H1 = ..."
template <class T> class optional;
template <class T> class optional<T&> {
    T* p_;
public:
    optional() : p_(nullptr) {}
    optional(none_t) : p_(nullptr) {}
    optional(T& r) : p_(&r) {}
    explicit operator bool() const { return p_ != nullptr; }
    T& operator*() const { return *p_; }
    T* operator->() const { return p_; }
};
}  // namespace boost

namespace gtsam {
typedef Eigen::MatrixXd Matrix;
typedef Eigen::VectorXd Vector;
typedef Eigen::Vector3d Vector3;
typedef std::uint64_t Key;
typedef std::function<std::string(Key)> KeyFormatter;
inline std::string stub_default_key_format(Key k) { return std::to_string(k); }
static const KeyFormatter DefaultKeyFormatter = &stub_default_key_format;

inline bool equal(double a, double b, double tol) { return std::fabs(a - b) <= tol; }
template <class A, class B>
inline bool equal(const Eigen::MatrixBase<A>& a, const Eigen::MatrixBase<B>& b, double tol) {
    if (a.rows() != b.rows() || a.cols() != b.cols()) return false;
    return ((a - b).array().abs() <= tol).all();
}

// "*H1 = *OptionalJacobian<15,15>(Hi);" : wraps a fixed-size matrix, dereferences to it
template <int R, int C> class OptionalJacobian {
    Eigen::Matrix<double, R, C>* m_;
public:
    OptionalJacobian(Eigen::Matrix<double, R, C>& m) : m_(&m) {}
    Eigen::Matrix<double, R, C>& operator*() { return *m_; }
};

namespace noiseModel {
struct Base {
    Matrix cov;
    virtual ~Base() {}
    virtual void print(const std::string& s) const { std::cout << s << "(stub Gaussian " << cov.rows() << "x" << cov.cols() << ")\n"; }
};
struct Gaussian : Base {
    typedef std::shared_ptr<Gaussian> shared_ptr;
    static shared_ptr Covariance(const Matrix& c) { shared_ptr p(new Gaussian()); p->cov = c; return p; }
};
}  // namespace noiseModel

class NonlinearFactor {
public:
    virtual ~NonlinearFactor() {}
};

template <class V1, class V2> class NoiseModelFactor2 : public NonlinearFactor {
protected:
    std::shared_ptr<noiseModel::Base> noiseModel_;
    Key k1_, k2_;
public:
    NoiseModelFactor2(const std::shared_ptr<noiseModel::Base>& nm, Key a, Key b) : noiseModel_(nm), k1_(a), k2_(b) {}
    Key key1() const { return k1_; }
    Key key2() const { return k2_; }
    bool equals(const NoiseModelFactor2& o, double tol) const {
        return k1_ == o.k1_ && k2_ == o.k2_ && gtsam::equal(noiseModel_->cov, o.noiseModel_->cov, tol);
    }
};

template <class T> struct traits;
namespace internal { template <class T> struct Manifold {}; }
}  // namespace gtsam
