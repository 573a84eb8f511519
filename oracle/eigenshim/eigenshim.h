// eigenshim.h -- TEST INFRASTRUCTURE ONLY (oracle/).  A small, eager stand-in for the subset of Eigen 3 that the reference's vendored g2o
// (SingleRobotScenario/Thirdparty/g2o) and Optimizer.cc / Converter.cc use, so that those files compile UNMODIFIED from where they lie and
// run as the reference's own object code (oracle/_ref/libref_g2o.so).  Eigen itself is an un-vendored, unpinned (>= 3.1.0) dependency of
// the reference and is absent from this image.
//
// What is restated here is Eigen's ARITHMETIC for the operations the path uses, each in the evaluation order Eigen documents / implements
// for small fixed sizes: coefficient-wise sums, matrix products as sum over k ascending, Quaternion(Matrix3) / toRotationMatrix /
// operator* / _transformVector (Eigen/src/Geometry/Quaternion.h), 2x2 / 3x3 / 4x4 inverse and determinant by cofactors
// (Eigen/src/LU/Inverse.h, Determinant.h), dense LLT / pivoted LDLT, a simplicial sparse LDL^T with a fill-reducing ordering.  No expression
// templates: every operator evaluates into a plain matrix (aliasing is therefore always safe).  Not a general Eigen replacement.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstring>
#include <cstdlib>
#include <iostream>
#include <limits>
#include <memory>
#include <vector>

#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW
#define EIGEN_MAKE_ALIGNED_OPERATOR_NEW_IF(x)
#define EIGEN_DEFINE_STL_VECTOR_SPECIALIZATION(...)
#define EIGEN_STRONG_INLINE inline
#define EIGEN_WORLD_VERSION 3
#define EIGEN_MAJOR_VERSION 2
#define EIGEN_MINOR_VERSION 0
#define EIGEN_VERSION_AT_LEAST(x, y, z) (EIGEN_WORLD_VERSION > x || (EIGEN_WORLD_VERSION >= x && (EIGEN_MAJOR_VERSION > y || (EIGEN_MAJOR_VERSION >= y && EIGEN_MINOR_VERSION >= z))))

namespace Eigen {

typedef std::ptrdiff_t DenseIndex;
typedef DenseIndex Index;
const int Dynamic = -1;
enum { ColMajor = 0, RowMajor = 0x1, AutoAlign = 0, DontAlign = 0x2 };
enum { Unaligned = 0, Aligned = 1 };
const unsigned int AlignedBit = 0x80;
enum { Lower = 0x1, Upper = 0x2, UnitDiag = 0x4, ZeroDiag = 0x8, UnitLower = UnitDiag | Lower, UnitUpper = UnitDiag | Upper, StrictlyLower = ZeroDiag | Lower, StrictlyUpper = ZeroDiag | Upper, SelfAdjoint = 0x10 };
enum ComputationInfo { Success = 0, NumericalIssue = 1, NoConvergence = 2, InvalidInput = 3 };
enum { ComputeEigenvectors = 0x80, EigenvaluesOnly = 0x40 };
enum TransformTraits { Isometry = 0x1, Affine = 0x2, AffineCompact = 0x10 | Affine, Projective = 0x20 };
inline void initParallel() {}

template <class T> class aligned_allocator : public std::allocator<T> {
  public:
    template <class U> struct rebind { typedef aligned_allocator<U> other; };
    aligned_allocator() {}
    template <class U> aligned_allocator(const aligned_allocator<U> &) {}
};

template <class S, int R, int C, int O = 0, int MR = R, int MC = C> class Matrix;
template <class Derived> class MatrixBase;
template <class X> class Transpose;
template <class X, int BR, int BC> class Block;
template <class X> class DiagView;
template <class X> class ArrayView;
template <class P, int MapOptions = Unaligned, class Stride = void> class Map;
template <class M> class LLT;
template <class M> class PartialPivLU;
template <class M> class LDLT;
template <class M> class SelfAdjointEigenSolver;

namespace internal {
template <class T> struct traits;
template <class S, int R, int C, int O, int MR, int MC> struct traits<Matrix<S, R, C, O, MR, MC> > { typedef S Scalar; enum { Rows = R, Cols = C }; };
template <class X> struct traits<const X> : traits<X> {};
template <class X> struct traits<Transpose<X> > { typedef typename traits<X>::Scalar Scalar; enum { Rows = traits<X>::Cols, Cols = traits<X>::Rows }; };
template <class X, int BR, int BC> struct traits<Block<X, BR, BC> > { typedef typename traits<X>::Scalar Scalar; enum { Rows = BR, Cols = BC }; };
template <class X> struct traits<DiagView<X> > { typedef typename traits<X>::Scalar Scalar; enum { Rows = (traits<X>::Rows == Dynamic || traits<X>::Cols == Dynamic) ? Dynamic : (traits<X>::Rows < traits<X>::Cols ? traits<X>::Rows : traits<X>::Cols), Cols = 1 }; };
template <class P, int O, class St> struct traits<Map<P, O, St> > : traits<P> {};
template <int A, int B> struct pick_dim { enum { value = (A == Dynamic) ? B : A }; };
}  // namespace internal

// ---- CRTP base: everything is expressed through rows() / cols() / coeff(i, j) [/ coeffRef(i, j)] of the derived class
template <class Derived> class MatrixBase {
  public:
    typedef typename internal::traits<Derived>::Scalar Scalar;
    enum { RowsAtCompileTime = internal::traits<Derived>::Rows, ColsAtCompileTime = internal::traits<Derived>::Cols,
           SizeAtCompileTime = (RowsAtCompileTime == Dynamic || ColsAtCompileTime == Dynamic) ? Dynamic : RowsAtCompileTime * ColsAtCompileTime,
           IsVectorAtCompileTime = RowsAtCompileTime == 1 || ColsAtCompileTime == 1, Flags = AlignedBit };
    typedef Matrix<Scalar, RowsAtCompileTime, ColsAtCompileTime> PlainObject;
    Derived &derived() { return *static_cast<Derived *>(this); }
    const Derived &derived() const { return *static_cast<const Derived *>(this); }
    Index rows() const { return derived().rows(); }
    Index cols() const { return derived().cols(); }
    Index size() const { return rows() * cols(); }
    Scalar coeff(Index i, Index j) const { return derived().coeff(i, j); }
    Scalar &coeffRef(Index i, Index j) { return derived().coeffRef(i, j); }
    Scalar coeff(Index i) const { return (ColsAtCompileTime == 1 || cols() == 1) ? derived().coeff(i, 0) : derived().coeff(0, i); }
    Scalar &coeffRef(Index i) { return (ColsAtCompileTime == 1 || cols() == 1) ? derived().coeffRef(i, 0) : derived().coeffRef(0, i); }
    Scalar operator()(Index i, Index j) const { return coeff(i, j); }
    Scalar &operator()(Index i, Index j) { return coeffRef(i, j); }
    Scalar operator()(Index i) const { return coeff(i); }
    Scalar &operator()(Index i) { return coeffRef(i); }
    Scalar operator[](Index i) const { return coeff(i); }
    Scalar &operator[](Index i) { return coeffRef(i); }
    Scalar x() const { return coeff(0); } Scalar y() const { return coeff(1); } Scalar z() const { return coeff(2); } Scalar w() const { return coeff(3); }
    Scalar &x() { return coeffRef(0); } Scalar &y() { return coeffRef(1); } Scalar &z() { return coeffRef(2); } Scalar &w() { return coeffRef(3); }
    PlainObject eval() const { return PlainObject(derived()); }
    Derived &noalias() { return derived(); }

    // ---- assignment family (the derived classes forward their operator= here)
    template <class O> Derived &assign(const MatrixBase<O> &o) {
        derived().resizeLike(o.rows(), o.cols());
        if ((const void *)&o == (const void *)this) return derived();
        for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) = o.coeff(i, j);
        return derived();
    }
    template <class O> Derived &operator+=(const MatrixBase<O> &o) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) += o.coeff(i, j); return derived(); }
    template <class O> Derived &operator-=(const MatrixBase<O> &o) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) -= o.coeff(i, j); return derived(); }
    Derived &operator*=(Scalar s) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) *= s; return derived(); }
    Derived &operator/=(Scalar s) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) /= s; return derived(); }
    template <class O> Derived &operator*=(const MatrixBase<O> &o) { PlainObject t = (*this) * o; return assign(t); }
    Derived &setZero() { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) = Scalar(0); return derived(); }
    Derived &setOnes() { return setConstant(Scalar(1)); }
    Derived &setConstant(Scalar v) { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) = v; return derived(); }
    void fill(Scalar v) { setConstant(v); }
    Derived &setIdentity() { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) coeffRef(i, j) = i == j ? Scalar(1) : Scalar(0); return derived(); }

    // ---- comma initialiser: m << a, b, c;  (row by row)
    struct CommaInit {
        Derived &m; Index k;
        CommaInit(Derived &mm, Scalar first) : m(mm), k(0) { put(first); }
        void put(Scalar v) { m.coeffRef(k / m.cols(), k % m.cols()) = v; k++; }
        CommaInit &operator,(Scalar v) { put(v); return *this; }
        Derived &finished() { return m; }
    };
    CommaInit operator<<(Scalar v) { return CommaInit(derived(), v); }

    // ---- views
    Transpose<Derived> transpose() { return Transpose<Derived>(derived()); }
    const Transpose<const Derived> transpose() const { return Transpose<const Derived>(derived()); }
    const Transpose<const Derived> adjoint() const { return Transpose<const Derived>(derived()); }
    template <int BR, int BC> Block<Derived, BR, BC> block(Index i, Index j) { return Block<Derived, BR, BC>(derived(), i, j, BR, BC); }
    template <int BR, int BC> const Block<const Derived, BR, BC> block(Index i, Index j) const { return Block<const Derived, BR, BC>(derived(), i, j, BR, BC); }
    Block<Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) { return Block<Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
    const Block<const Derived, Dynamic, Dynamic> block(Index i, Index j, Index r, Index c) const { return Block<const Derived, Dynamic, Dynamic>(derived(), i, j, r, c); }
    template <int BR, int BC> Block<Derived, BR, BC> topLeftCorner() { return block<BR, BC>(0, 0); }
    template <int BR, int BC> const Block<const Derived, BR, BC> topLeftCorner() const { return block<BR, BC>(0, 0); }
    template <int BR, int BC> Block<Derived, BR, BC> topRightCorner() { return block<BR, BC>(0, cols() - BC); }
    template <int BR, int BC> const Block<const Derived, BR, BC> topRightCorner() const { return block<BR, BC>(0, cols() - BC); }
    Block<Derived, RowsAtCompileTime, 1> col(Index j) { return Block<Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
    const Block<const Derived, RowsAtCompileTime, 1> col(Index j) const { return Block<const Derived, RowsAtCompileTime, 1>(derived(), 0, j, rows(), 1); }
    Block<Derived, 1, ColsAtCompileTime> row(Index i) { return Block<Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
    const Block<const Derived, 1, ColsAtCompileTime> row(Index i) const { return Block<const Derived, 1, ColsAtCompileTime>(derived(), i, 0, 1, cols()); }
    // vector segments (column or row vectors)
    template <int N> Block<Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)> segment(Index s) {
        return Block<Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)>(derived(), ColsAtCompileTime == 1 ? s : 0, ColsAtCompileTime == 1 ? 0 : s, ColsAtCompileTime == 1 ? N : 1, ColsAtCompileTime == 1 ? 1 : N); }
    template <int N> const Block<const Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)> segment(Index s) const {
        return Block<const Derived, (ColsAtCompileTime == 1 ? N : 1), (ColsAtCompileTime == 1 ? 1 : N)>(derived(), ColsAtCompileTime == 1 ? s : 0, ColsAtCompileTime == 1 ? 0 : s, ColsAtCompileTime == 1 ? N : 1, ColsAtCompileTime == 1 ? 1 : N); }
    Block<Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)> segment(Index s, Index n) {
        return Block<Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)>(derived(), ColsAtCompileTime == 1 ? s : 0, ColsAtCompileTime == 1 ? 0 : s, ColsAtCompileTime == 1 ? n : 1, ColsAtCompileTime == 1 ? 1 : n); }
    const Block<const Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)> segment(Index s, Index n) const {
        return Block<const Derived, (ColsAtCompileTime == 1 ? Dynamic : 1), (ColsAtCompileTime == 1 ? 1 : Dynamic)>(derived(), ColsAtCompileTime == 1 ? s : 0, ColsAtCompileTime == 1 ? 0 : s, ColsAtCompileTime == 1 ? n : 1, ColsAtCompileTime == 1 ? 1 : n); }
    template <int N> auto head() -> decltype(this->template segment<N>(0)) { return this->template segment<N>(0); }
    template <int N> auto head() const -> decltype(this->template segment<N>(0)) { return this->template segment<N>(0); }
    template <int N> auto tail() -> decltype(this->template segment<N>(0)) { return this->template segment<N>(size() - N); }
    template <int N> auto tail() const -> decltype(this->template segment<N>(0)) { return this->template segment<N>(size() - N); }
    auto head(Index n) -> decltype(this->segment(0, n)) { return segment(0, n); }
    auto head(Index n) const -> decltype(this->segment(0, n)) { return segment(0, n); }
    auto tail(Index n) -> decltype(this->segment(0, n)) { return segment(size() - n, n); }
    auto tail(Index n) const -> decltype(this->segment(0, n)) { return segment(size() - n, n); }
    DiagView<Derived> diagonal() { return DiagView<Derived>(derived()); }
    const DiagView<const Derived> diagonal() const { return DiagView<const Derived>(derived()); }
    ArrayView<Derived> array() { return ArrayView<Derived>(derived()); }
    const Derived &matrix() const { return derived(); }

    // ---- reductions
    Scalar squaredNorm() const { Scalar s = 0; for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) s += coeff(i, j) * coeff(i, j); return s; }
    Scalar norm() const { return std::sqrt(squaredNorm()); }
    Scalar sum() const { Scalar s = 0; for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) s += coeff(i, j); return s; }
    Scalar trace() const { Scalar s = 0; for (Index i = 0; i < std::min(rows(), cols()); i++) s += coeff(i, i); return s; }
    Scalar maxCoeff() const { Scalar m = coeff(0, 0); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m = std::max(m, coeff(i, j)); return m; }
    Scalar minCoeff() const { Scalar m = coeff(0, 0); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) m = std::min(m, coeff(i, j)); return m; }
    template <class O> Scalar dot(const MatrixBase<O> &o) const { Scalar s = 0; for (Index i = 0; i < size(); i++) s += coeff(i) * o.coeff(i); return s; }
    template <class O> Matrix<Scalar, 3, 1> cross(const MatrixBase<O> &o) const;
    void normalize() { const Scalar n = norm(); *this /= n; }
    PlainObject normalized() const { PlainObject r(derived()); r /= norm(); return r; }
    PlainObject cwiseAbs() const { PlainObject r(derived()); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) r.coeffRef(i, j) = std::abs(coeff(i, j)); return r; }
    template <class O> PlainObject cwiseProduct(const MatrixBase<O> &o) const { PlainObject r(derived()); for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) r.coeffRef(i, j) = coeff(i, j) * o.coeff(i, j); return r; }
    bool allFinite() const { for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) if (!std::isfinite(coeff(i, j))) return false; return true; }
    template <class O> bool isApprox(const MatrixBase<O> &o, Scalar prec = 1e-12) const { Scalar d = 0; for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) { const Scalar e = coeff(i, j) - o.coeff(i, j); d += e * e; } return d <= prec * prec * std::min(squaredNorm(), o.squaredNorm()); }
    template <class NewScalar> Matrix<NewScalar, RowsAtCompileTime, ColsAtCompileTime> cast() const {
        Matrix<NewScalar, RowsAtCompileTime, ColsAtCompileTime> r; r.resizeLike(rows(), cols());
        for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) r.coeffRef(i, j) = NewScalar(coeff(i, j));
        return r;
    }

    template <class O> bool operator==(const MatrixBase<O> &o) const { if (rows() != o.rows() || cols() != o.cols()) return false; for (Index j = 0; j < cols(); j++) for (Index i = 0; i < rows(); i++) if (!(coeff(i, j) == o.coeff(i, j))) return false; return true; }
    template <class O> bool operator!=(const MatrixBase<O> &o) const { return !(*this == o); }

    // ---- dense linear algebra (defined below)
    PartialPivLU<PlainObject> lu() const;
    PartialPivLU<PlainObject> partialPivLu() const;
    Scalar determinant() const;
    PlainObject inverse() const;
    LLT<PlainObject> llt() const;
    LDLT<PlainObject> ldlt() const;

    // ---- statics
    static PlainObject Zero() { PlainObject r; r.setZero(); return r; }
    static PlainObject Zero(Index n) { PlainObject r; r.resizeLike(ColsAtCompileTime == 1 ? n : 1, ColsAtCompileTime == 1 ? 1 : n); r.setZero(); return r; }
    static PlainObject Zero(Index r_, Index c_) { PlainObject r; r.resizeLike(r_, c_); r.setZero(); return r; }
    static PlainObject Ones() { PlainObject r; r.setOnes(); return r; }
    static PlainObject Constant(Scalar v) { PlainObject r; r.setConstant(v); return r; }
    static PlainObject Identity() { PlainObject r; r.setIdentity(); return r; }
    static PlainObject Identity(Index r_, Index c_) { PlainObject r; r.resizeLike(r_, c_); r.setIdentity(); return r; }
    static PlainObject Random() { PlainObject r; for (Index j = 0; j < r.cols(); j++) for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = Scalar(2) * Scalar(std::rand()) / Scalar(RAND_MAX) - Scalar(1); return r; }
};

// ---- plain storage
template <class S, int R, int C, int O, int MR, int MC> class Matrix : public MatrixBase<Matrix<S, R, C, O, MR, MC> > {
    typedef MatrixBase<Matrix> Base;
    enum { Fixed = (R != Dynamic && C != Dynamic) };
    S fix_[Fixed ? (R * C > 0 ? R * C : 1) : 1];
    std::vector<S> dyn_;
    Index r_, c_;
  public:
    typedef S Scalar;
    typedef Map<Matrix, Unaligned> MapType;
    typedef const Map<const Matrix, Unaligned> ConstMapType;
    typedef Map<Matrix, Aligned> AlignedMapType;
    Matrix() : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) {}
    Matrix(const Matrix &o) : dyn_(o.dyn_), r_(o.r_), c_(o.c_) { if (Fixed) std::memcpy(fix_, o.fix_, sizeof fix_); }
    explicit Matrix(Index n) : r_(R == Dynamic ? (C == 1 || R == Dynamic ? n : 0) : R), c_(C == Dynamic ? (R == 1 ? n : (R == Dynamic ? 1 : n)) : C) {
        if (R == Dynamic && C == Dynamic) { r_ = n; c_ = 1; }
        if (!Fixed) dyn_.assign((size_t)(r_ * c_), S(0));
        else if (R * C == 1) fix_[0] = S(n);
    }
    // (rows, cols) for every size but the fixed 2-vector, where it is (x, y) -- like Eigen
    template <class T0, class T1> Matrix(const T0 &a, const T1 &b) : r_(R == Dynamic ? (Index)a : R), c_(C == Dynamic ? (Index)b : C)
    {
        if (!Fixed) dyn_.assign((size_t)(r_ * c_), S(0));
        else if (R * C == 2) { fix_[0] = S(a); fix_[1] = S(b); }
    }
    Matrix(const S &x, const S &y, const S &z) : r_(R), c_(C) { fix_[0] = x; fix_[1] = y; fix_[2] = z; }
    Matrix(const S &x, const S &y, const S &z, const S &w) : r_(R), c_(C) { fix_[0] = x; fix_[1] = y; fix_[2] = z; fix_[3] = w; }
    explicit Matrix(const S *d) : r_(R), c_(C) { static_assert(Fixed, "data constructor on a dynamic matrix"); std::memcpy(fix_, d, sizeof(S) * R * C); }
    template <class Od> Matrix(const MatrixBase<Od> &o) : r_(R == Dynamic ? 0 : R), c_(C == Dynamic ? 0 : C) { Base::assign(o); }
    Matrix &operator=(const Matrix &o) { if (this != &o) { r_ = o.r_; c_ = o.c_; dyn_ = o.dyn_; if (Fixed) std::memcpy(fix_, o.fix_, sizeof fix_); } return *this; }
    template <class Od> Matrix &operator=(const MatrixBase<Od> &o) { return Base::assign(o); }
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    S *data() { return Fixed ? fix_ : dyn_.data(); }
    const S *data() const { return Fixed ? fix_ : dyn_.data(); }
    S coeff(Index i, Index j) const { return (O & RowMajor) ? data()[i * c_ + j] : data()[j * r_ + i]; }
    S &coeffRef(Index i, Index j) { return (O & RowMajor) ? data()[i * c_ + j] : data()[j * r_ + i]; }
    using Base::coeff; using Base::coeffRef;
    void resize(Index r, Index c) { assert((R == Dynamic || r == R) && (C == Dynamic || c == C)); if (!Fixed && (r != r_ || c != c_)) { r_ = r; c_ = c; dyn_.assign((size_t)(r * c), S(0)); } }
    void resize(Index n) { if (C == 1 || (R == Dynamic && C == Dynamic)) resize(n, C == Dynamic ? 1 : C); else resize(R == Dynamic ? 1 : R, n); }
    void resizeLike(Index r, Index c) { resize(r, c); }
    void conservativeResize(Index r, Index c) { Matrix t(*this); const Index orr = r_, oc = c_; resize(r, c); Base::setZero(); for (Index j = 0; j < std::min(oc, c); j++) for (Index i = 0; i < std::min(orr, r); i++) coeffRef(i, j) = t.coeff(i, j); }
    void conservativeResize(Index n) { if (C == 1) conservativeResize(n, 1); else conservativeResize(1, n); }
    void swap(Matrix &o) { std::swap(*this, o); }
};

template <class X> class Transpose : public MatrixBase<Transpose<X> > {
    X &x_;
  public:
    typedef typename internal::traits<X>::Scalar Scalar;
    explicit Transpose(X &x) : x_(x) {}
    Index rows() const { return x_.cols(); }
    Index cols() const { return x_.rows(); }
    Scalar coeff(Index i, Index j) const { return x_.coeff(j, i); }
    Scalar &coeffRef(Index i, Index j) { return const_cast<typename std::remove_const<X>::type &>(x_).coeffRef(j, i); }
    using MatrixBase<Transpose>::coeff; using MatrixBase<Transpose>::coeffRef;
    void resizeLike(Index r, Index c) const { assert(r == rows() && c == cols()); (void)r; (void)c; }
    template <class Od> Transpose &operator=(const MatrixBase<Od> &o) { typename MatrixBase<Od>::PlainObject t(o.derived()); return MatrixBase<Transpose>::assign(t); }
};

template <class X, int BR, int BC> class Block : public MatrixBase<Block<X, BR, BC> > {
    X &x_; Index i0_, j0_, r_, c_;
  public:
    typedef typename internal::traits<X>::Scalar Scalar;
    Block(X &x, Index i0, Index j0, Index r, Index c) : x_(x), i0_(i0), j0_(j0), r_(r), c_(c) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    Scalar coeff(Index i, Index j) const { return x_.coeff(i0_ + i, j0_ + j); }
    Scalar &coeffRef(Index i, Index j) { return const_cast<typename std::remove_const<X>::type &>(x_).coeffRef(i0_ + i, j0_ + j); }
    using MatrixBase<Block>::coeff; using MatrixBase<Block>::coeffRef;
    void resizeLike(Index r, Index c) const { assert(r == r_ && c == c_); (void)r; (void)c; }
    Block &operator=(const Block &o) { typename MatrixBase<Block>::PlainObject t(o); return MatrixBase<Block>::assign(t); }
    template <class Od> Block &operator=(const MatrixBase<Od> &o) { typename MatrixBase<Od>::PlainObject t(o.derived()); return MatrixBase<Block>::assign(t); }
};

template <class X> class DiagView : public MatrixBase<DiagView<X> > {
    X &x_;
  public:
    typedef typename internal::traits<X>::Scalar Scalar;
    explicit DiagView(X &x) : x_(x) {}
    Index rows() const { return std::min(x_.rows(), x_.cols()); }
    Index cols() const { return 1; }
    Scalar coeff(Index i, Index) const { return x_.coeff(i, i); }
    Scalar &coeffRef(Index i, Index) { return const_cast<typename std::remove_const<X>::type &>(x_).coeffRef(i, i); }
    using MatrixBase<DiagView>::coeff; using MatrixBase<DiagView>::coeffRef;
    void resizeLike(Index, Index) const {}
    template <class Od> DiagView &operator=(const MatrixBase<Od> &o) { return MatrixBase<DiagView>::assign(o); }
};

template <class X> class ArrayView {
    X &x_;
  public:
    typedef typename internal::traits<X>::Scalar Scalar;
    explicit ArrayView(X &x) : x_(x) {}
    ArrayView &operator+=(Scalar s) { for (Index j = 0; j < x_.cols(); j++) for (Index i = 0; i < x_.rows(); i++) x_.coeffRef(i, j) += s; return *this; }
    ArrayView &operator-=(Scalar s) { return *this += -s; }
    ArrayView &operator*=(Scalar s) { for (Index j = 0; j < x_.cols(); j++) for (Index i = 0; i < x_.rows(); i++) x_.coeffRef(i, j) *= s; return *this; }
};

// Map: a view on caller-owned memory, re-seatable by placement new (g2o's mapHessianMemory)
template <class P, int MapOptions, class Stride> class Map : public MatrixBase<Map<P, MapOptions, Stride> > {
    typedef typename std::remove_const<P>::type Plain;
    enum { R = internal::traits<Plain>::Rows, C = internal::traits<Plain>::Cols };
  public:
    typedef typename internal::traits<Plain>::Scalar Scalar;
    typedef typename std::conditional<std::is_const<P>::value, const Scalar *, Scalar *>::type Pointer;
  private:
    Pointer d_; Index r_, c_;
  public:
    explicit Map(Pointer d) : d_(d), r_(R), c_(C) {}
    Map(Pointer d, Index n) : d_(d), r_(C == 1 ? n : (R == Dynamic && C == Dynamic ? n : R)), c_(C == 1 ? 1 : (R == Dynamic && C == Dynamic ? 1 : n)) {}
    Map(Pointer d, Index r, Index c) : d_(d), r_(r), c_(c) {}
    Index rows() const { return r_; }
    Index cols() const { return c_; }
    const Scalar *data() const { return d_; }
    Pointer data() { return d_; }
    Scalar coeff(Index i, Index j) const { return d_[j * r_ + i]; }
    Scalar &coeffRef(Index i, Index j) { return const_cast<Scalar *>(d_)[j * r_ + i]; }
    using MatrixBase<Map>::coeff; using MatrixBase<Map>::coeffRef;
    void resizeLike(Index r, Index c) const { assert(r == r_ && c == c_); (void)r; (void)c; }
    Map &operator=(const Map &o) { Plain t(o); return MatrixBase<Map>::assign(t); }
    template <class Od> Map &operator=(const MatrixBase<Od> &o) { typename MatrixBase<Od>::PlainObject t(o.derived()); return MatrixBase<Map>::assign(t); }
};

// ---- operators (all eager)
template <class A, class B> typename MatrixBase<A>::PlainObject operator+(const MatrixBase<A> &a, const MatrixBase<B> &b) { typename MatrixBase<A>::PlainObject r(a.derived()); r += b; return r; }
template <class A, class B> typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A> &a, const MatrixBase<B> &b) { typename MatrixBase<A>::PlainObject r(a.derived()); r -= b; return r; }
template <class A> typename MatrixBase<A>::PlainObject operator-(const MatrixBase<A> &a) { typename MatrixBase<A>::PlainObject r(a.derived()); for (Index j = 0; j < r.cols(); j++) for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = -r.coeff(i, j); return r; }
template <class A> typename MatrixBase<A>::PlainObject operator*(const MatrixBase<A> &a, typename MatrixBase<A>::Scalar s) { typename MatrixBase<A>::PlainObject r(a.derived()); r *= s; return r; }
template <class A> typename MatrixBase<A>::PlainObject operator*(typename MatrixBase<A>::Scalar s, const MatrixBase<A> &a) { typename MatrixBase<A>::PlainObject r(a.derived()); for (Index j = 0; j < r.cols(); j++) for (Index i = 0; i < r.rows(); i++) r.coeffRef(i, j) = s * r.coeff(i, j); return r; }
template <class A> typename MatrixBase<A>::PlainObject operator/(const MatrixBase<A> &a, typename MatrixBase<A>::Scalar s) { typename MatrixBase<A>::PlainObject r(a.derived()); r /= s; return r; }
template <class A, class B>
Matrix<typename MatrixBase<A>::Scalar, MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::ColsAtCompileTime> operator*(const MatrixBase<A> &a, const MatrixBase<B> &b)
{
    typedef typename MatrixBase<A>::Scalar S;
    Matrix<S, MatrixBase<A>::RowsAtCompileTime, MatrixBase<B>::ColsAtCompileTime> r;
    r.resizeLike(a.rows(), b.cols());
    assert(a.cols() == b.rows());
    const Index n = a.cols();
    for (Index j = 0; j < b.cols(); j++)
        for (Index i = 0; i < a.rows(); i++) {
            S s = n > 0 ? a.coeff(i, 0) * b.coeff(0, j) : S(0);
            for (Index k = 1; k < n; k++) s += a.coeff(i, k) * b.coeff(k, j);
            r.coeffRef(i, j) = s;
        }
    return r;
}
template <class A> std::ostream &operator<<(std::ostream &os, const MatrixBase<A> &a)
{
    for (Index i = 0; i < a.rows(); i++) { for (Index j = 0; j < a.cols(); j++) os << (j ? " " : "") << a.coeff(i, j); if (i + 1 < a.rows()) os << "\n"; }
    return os;
}
template <class D> template <class O> Matrix<typename MatrixBase<D>::Scalar, 3, 1> MatrixBase<D>::cross(const MatrixBase<O> &o) const
{
    return Matrix<Scalar, 3, 1>(coeff(1) * o.coeff(2) - coeff(2) * o.coeff(1), coeff(2) * o.coeff(0) - coeff(0) * o.coeff(2), coeff(0) * o.coeff(1) - coeff(1) * o.coeff(0));
}

typedef Matrix<double, 2, 1> Vector2d; typedef Matrix<double, 3, 1> Vector3d; typedef Matrix<double, 4, 1> Vector4d;
typedef Matrix<double, Dynamic, 1> VectorXd; typedef Matrix<double, Dynamic, Dynamic> MatrixXd;
typedef Matrix<double, 2, 2> Matrix2d; typedef Matrix<double, 3, 3> Matrix3d; typedef Matrix<double, 4, 4> Matrix4d;
typedef Matrix<float, 2, 1> Vector2f; typedef Matrix<float, 3, 1> Vector3f; typedef Matrix<float, 3, 3> Matrix3f; typedef Matrix<float, 4, 4> Matrix4f;
typedef Matrix<int, 2, 1> Vector2i; typedef Matrix<int, 3, 1> Vector3i; typedef Matrix<int, Dynamic, 1> VectorXi;
typedef Matrix<double, 1, Dynamic> RowVectorXd;

// ---- determinant / inverse: cofactor formulas up to 4x4 (Eigen/src/LU/Determinant.h, Inverse.h), partial-pivot LU beyond
namespace internal {
template <class M> typename M::Scalar det3(const M &m, int a, int b, int c) { return m.coeff(0, a) * (m.coeff(1, b) * m.coeff(2, c) - m.coeff(1, c) * m.coeff(2, b)); }
template <class S> S lu_det_inv(std::vector<S> a, Index n, std::vector<S> *inv)
{
    std::vector<S> b;
    if (inv) { b.assign((size_t)(n * n), S(0)); for (Index i = 0; i < n; i++) b[i * n + i] = S(1); }
    S det = S(1);
    for (Index k = 0; k < n; k++) {
        Index p = k;
        for (Index i = k + 1; i < n; i++) if (std::abs(a[i * n + k]) > std::abs(a[p * n + k])) p = i;
        if (p != k) { for (Index j = 0; j < n; j++) { std::swap(a[k * n + j], a[p * n + j]); if (inv) std::swap(b[k * n + j], b[p * n + j]); } det = -det; }
        det *= a[k * n + k];
        if (a[k * n + k] == S(0)) continue;
        for (Index i = k + 1; i < n; i++) {
            const S f = a[i * n + k] / a[k * n + k];
            for (Index j = k; j < n; j++) a[i * n + j] -= f * a[k * n + j];
            if (inv) for (Index j = 0; j < n; j++) b[i * n + j] -= f * b[k * n + j];
        }
    }
    if (inv) {
        for (Index j = 0; j < n; j++)
            for (Index i = n - 1; i >= 0; i--) { S s = b[i * n + j]; for (Index k = i + 1; k < n; k++) s -= a[i * n + k] * b[k * n + j]; b[i * n + j] = s / a[i * n + i]; }
        *inv = b;
    }
    return det;
}
}  // namespace internal
template <class D> typename MatrixBase<D>::Scalar MatrixBase<D>::determinant() const
{
    const Index n = rows();
    if (n == 1) return coeff(0, 0);
    if (n == 2) return coeff(0, 0) * coeff(1, 1) - coeff(1, 0) * coeff(0, 1);
    if (n == 3) return internal::det3(derived(), 0, 1, 2) - internal::det3(derived(), 1, 0, 2) + internal::det3(derived(), 2, 0, 1);
    std::vector<Scalar> a((size_t)(n * n));
    for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) a[i * n + j] = coeff(i, j);
    return internal::lu_det_inv<Scalar>(a, n, nullptr);
}
template <class D> typename MatrixBase<D>::PlainObject MatrixBase<D>::inverse() const
{
    const Index n = rows();
    PlainObject r; r.resizeLike(n, n);
    if (n == 1) { r.coeffRef(0, 0) = Scalar(1) / coeff(0, 0); return r; }
    if (n == 2) {
        const Scalar invdet = Scalar(1) / determinant();
        r.coeffRef(0, 0) = coeff(1, 1) * invdet; r.coeffRef(1, 0) = -coeff(1, 0) * invdet; r.coeffRef(0, 1) = -coeff(0, 1) * invdet; r.coeffRef(1, 1) = coeff(0, 0) * invdet;
        return r;
    }
    if (n == 3) {
        // cofactor_3x3<i, j>: minor with rows (i+1)%3, (i+2)%3 and columns (j+1)%3, (j+2)%3
        auto cof = [&](int i, int j) { const int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3; return coeff(i1, j1) * coeff(i2, j2) - coeff(i1, j2) * coeff(i2, j1); };
        const Scalar c00 = cof(0, 0), c10 = cof(1, 0), c20 = cof(2, 0);
        const Scalar det = (c00 * coeff(0, 0) + c10 * coeff(1, 0)) + c20 * coeff(2, 0);
        const Scalar invdet = Scalar(1) / det;
        r.coeffRef(0, 0) = c00 * invdet; r.coeffRef(0, 1) = c10 * invdet; r.coeffRef(0, 2) = c20 * invdet;
        r.coeffRef(1, 0) = cof(0, 1) * invdet; r.coeffRef(1, 1) = cof(1, 1) * invdet; r.coeffRef(1, 2) = cof(2, 1) * invdet;
        r.coeffRef(2, 0) = cof(0, 2) * invdet; r.coeffRef(2, 1) = cof(1, 2) * invdet; r.coeffRef(2, 2) = cof(2, 2) * invdet;
        return r;
    }
    std::vector<Scalar> a((size_t)(n * n)), b;
    for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) a[i * n + j] = coeff(i, j);
    internal::lu_det_inv<Scalar>(a, n, &b);
    for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) r.coeffRef(i, j) = b[i * n + j];
    return r;
}

// ---- dense Cholesky: LLT (lower), LDLT with diagonal pivoting (Eigen/src/Cholesky/LDLT.h: pivot = largest remaining |diagonal|)
template <class M> class LLT {
    typedef typename M::Scalar S;
    MatrixXd L_; bool ok_;
  public:
    LLT() : ok_(false) {}
    template <class A> explicit LLT(const MatrixBase<A> &a) { compute(a); }
    template <class A> LLT &compute(const MatrixBase<A> &a)
    {
        const Index n = a.rows(); L_ = MatrixXd::Zero(n, n); ok_ = true;
        for (Index j = 0; j < n; j++) {
            double d = a.coeff(j, j);
            for (Index k = 0; k < j; k++) d -= L_(j, k) * L_(j, k);
            if (!(d > 0)) { ok_ = false; d = std::abs(d) > 0 ? std::abs(d) : 1; }
            const double l = std::sqrt(d); L_(j, j) = l;
            for (Index i = j + 1; i < n; i++) { double s = a.coeff(i, j); for (Index k = 0; k < j; k++) s -= L_(i, k) * L_(j, k); L_(i, j) = s / l; }
        }
        return *this;
    }
    ComputationInfo info() const { return ok_ ? Success : NumericalIssue; }
    const MatrixXd &matrixL() const { return L_; }
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B> &b) const
    {
        typename MatrixBase<B>::PlainObject x(b.derived());
        const Index n = L_.rows();
        for (Index c = 0; c < x.cols(); c++) {
            for (Index i = 0; i < n; i++) { double s = x.coeff(i, c); for (Index k = 0; k < i; k++) s -= L_(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s / L_(i, i); }
            for (Index i = n - 1; i >= 0; i--) { double s = x.coeff(i, c); for (Index k = i + 1; k < n; k++) s -= L_(k, i) * x.coeff(k, c); x.coeffRef(i, c) = s / L_(i, i); }
        }
        return x;
    }
};
template <class M> class LDLT {
    MatrixXd L_; std::vector<double> D_; std::vector<Index> perm_; bool pos_, neg_;
  public:
    LDLT() : pos_(true), neg_(true) {}
    template <class A> explicit LDLT(const MatrixBase<A> &a) { compute(a); }
    template <class A> LDLT &compute(const MatrixBase<A> &a)
    {
        const Index n = a.rows();
        MatrixXd W(n, n);
        for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) W(i, j) = a.coeff(std::max(i, j), std::min(i, j));   // the lower triangle is referenced
        perm_.resize((size_t)n); for (Index i = 0; i < n; i++) perm_[(size_t)i] = i;
        L_ = MatrixXd::Identity(n, n); D_.assign((size_t)n, 0.0); pos_ = neg_ = true;
        for (Index k = 0; k < n; k++) {
            Index p = k;
            for (Index i = k + 1; i < n; i++) if (std::abs(W(i, i)) > std::abs(W(p, p))) p = i;
            if (p != k) {
                for (Index j = 0; j < n; j++) std::swap(W(k, j), W(p, j));
                for (Index i = 0; i < n; i++) std::swap(W(i, k), W(i, p));
                for (Index j = 0; j < k; j++) std::swap(L_(k, j), L_(p, j));
                std::swap(perm_[(size_t)k], perm_[(size_t)p]);
            }
            const double d = W(k, k); D_[(size_t)k] = d;
            if (d > 0) neg_ = false; else if (d < 0) pos_ = false;
            if (d == 0) continue;
            for (Index i = k + 1; i < n; i++) L_(i, k) = W(i, k) / d;
            for (Index i = k + 1; i < n; i++) for (Index j = k + 1; j < n; j++) W(i, j) -= L_(i, k) * W(k, j);
        }
        return *this;
    }
    bool isPositive() const { return pos_; }
    bool isNegative() const { return neg_; }
    ComputationInfo info() const { return Success; }
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B> &b) const
    {
        const Index n = L_.rows();
        typename MatrixBase<B>::PlainObject x(b.derived());
        for (Index c = 0; c < x.cols(); c++) {
            std::vector<double> y((size_t)n);
            for (Index i = 0; i < n; i++) y[(size_t)i] = b.coeff(perm_[(size_t)i], c);
            for (Index i = 0; i < n; i++) { double s = y[(size_t)i]; for (Index k = 0; k < i; k++) s -= L_(i, k) * y[(size_t)k]; y[(size_t)i] = s; }
            for (Index i = 0; i < n; i++) y[(size_t)i] = D_[(size_t)i] != 0 ? y[(size_t)i] / D_[(size_t)i] : 0.0;
            for (Index i = n - 1; i >= 0; i--) { double s = y[(size_t)i]; for (Index k = i + 1; k < n; k++) s -= L_(k, i) * y[(size_t)k]; y[(size_t)i] = s; }
            for (Index i = 0; i < n; i++) x.coeffRef(perm_[(size_t)i], c) = y[(size_t)i];
        }
        return x;
    }
};
// LU with partial (row) pivoting, Eigen/src/LU/PartialPivLU.h: pivot = largest |entry| of the column, first one on ties
template <class M> class PartialPivLU {
    MatrixXd lu_; std::vector<Index> piv_;
  public:
    template <class A> explicit PartialPivLU(const MatrixBase<A> &a) : lu_(a.derived())
    {
        const Index n = lu_.rows(); piv_.resize((size_t)n);
        for (Index k = 0; k < n; k++) {
            Index p = k;
            for (Index i = k + 1; i < n; i++) if (std::abs(lu_(i, k)) > std::abs(lu_(p, k))) p = i;
            piv_[(size_t)k] = p;
            if (p != k) for (Index j = 0; j < n; j++) std::swap(lu_(k, j), lu_(p, j));
            if (lu_(k, k) == 0) continue;
            for (Index i = k + 1; i < n; i++) { lu_(i, k) /= lu_(k, k); for (Index j = k + 1; j < n; j++) lu_(i, j) -= lu_(i, k) * lu_(k, j); }
        }
    }
    template <class B> typename MatrixBase<B>::PlainObject solve(const MatrixBase<B> &b) const
    {
        typename MatrixBase<B>::PlainObject x(b.derived());
        const Index n = lu_.rows();
        for (Index c = 0; c < x.cols(); c++) {
            for (Index k = 0; k < n; k++) if (piv_[(size_t)k] != k) std::swap(x.coeffRef(k, c), x.coeffRef(piv_[(size_t)k], c));
            for (Index i = 0; i < n; i++) { double s = x.coeff(i, c); for (Index k = 0; k < i; k++) s -= lu_(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s; }
            for (Index i = n - 1; i >= 0; i--) { double s = x.coeff(i, c); for (Index k = i + 1; k < n; k++) s -= lu_(i, k) * x.coeff(k, c); x.coeffRef(i, c) = s / lu_(i, i); }
        }
        return x;
    }
};
template <class D> PartialPivLU<typename MatrixBase<D>::PlainObject> MatrixBase<D>::lu() const { return PartialPivLU<PlainObject>(derived()); }
template <class D> PartialPivLU<typename MatrixBase<D>::PlainObject> MatrixBase<D>::partialPivLu() const { return PartialPivLU<PlainObject>(derived()); }
template <class D> LLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::llt() const { return LLT<PlainObject>(derived()); }
template <class D> LDLT<typename MatrixBase<D>::PlainObject> MatrixBase<D>::ldlt() const { return LDLT<PlainObject>(derived()); }

// symmetric eigenvalues (cyclic Jacobi): only the information-matrix sanity check of OptimizableGraph::verifyInformationMatrices uses it
template <class M> class SelfAdjointEigenSolver {
    VectorXd ev_;
  public:
    SelfAdjointEigenSolver() {}
    template <class A> SelfAdjointEigenSolver &compute(const MatrixBase<A> &a, int = ComputeEigenvectors)
    {
        const Index n = a.rows(); MatrixXd W(n, n);
        for (Index i = 0; i < n; i++) for (Index j = 0; j < n; j++) W(i, j) = a.coeff(std::max(i, j), std::min(i, j));
        for (int sweep = 0; sweep < 60; sweep++) {
            double off = 0; for (Index i = 0; i < n; i++) for (Index j = 0; j < i; j++) off += W(i, j) * W(i, j);
            if (off < 1e-300) break;
            for (Index p = 0; p < n; p++) for (Index q = p + 1; q < n; q++) {
                if (W(p, q) == 0) continue;
                const double th = (W(q, q) - W(p, p)) / (2 * W(p, q)), t = (th >= 0 ? 1 : -1) / (std::abs(th) + std::sqrt(th * th + 1)), c = 1 / std::sqrt(t * t + 1), s = t * c;
                for (Index k = 0; k < n; k++) { const double a1 = W(k, p), a2 = W(k, q); W(k, p) = c * a1 - s * a2; W(k, q) = s * a1 + c * a2; }
                for (Index k = 0; k < n; k++) { const double a1 = W(p, k), a2 = W(q, k); W(p, k) = c * a1 - s * a2; W(q, k) = s * a1 + c * a2; }
            }
        }
        ev_.resize(n); for (Index i = 0; i < n; i++) ev_(i) = W(i, i);
        std::sort(ev_.data(), ev_.data() + n);
        return *this;
    }
    const VectorXd &eigenvalues() const { return ev_; }
};

// ---- Geometry: Quaternion (coefficients stored x y z w), AngleAxis, a minimal Transform
template <class S> class AngleAxis;
template <class S, int Opt = 0> class Quaternion {
    Matrix<S, 4, 1> c_;
  public:
    typedef S Scalar;
    typedef Matrix<S, 3, 1> Vector3; typedef Matrix<S, 3, 3> Matrix3;
    Quaternion() {}
    Quaternion(const S &w, const S &x, const S &y, const S &z) : c_(x, y, z, w) {}
    explicit Quaternion(const S *d) : c_(d) {}
    template <class Dv> explicit Quaternion(const MatrixBase<Dv> &m) { *this = m; }
    explicit Quaternion(const AngleAxis<S> &aa);
    S x() const { return c_[0]; } S y() const { return c_[1]; } S z() const { return c_[2]; } S w() const { return c_[3]; }
    S &x() { return c_[0]; } S &y() { return c_[1]; } S &z() { return c_[2]; } S &w() { return c_[3]; }
    const Matrix<S, 4, 1> &coeffs() const { return c_; }
    Matrix<S, 4, 1> &coeffs() { return c_; }
    Vector3 vec() const { return Vector3(c_[0], c_[1], c_[2]); }
    static Quaternion Identity() { return Quaternion(S(1), S(0), S(0), S(0)); }
    Quaternion &setIdentity() { c_ = Matrix<S, 4, 1>(S(0), S(0), S(0), S(1)); return *this; }
    S squaredNorm() const { return c_.squaredNorm(); }
    S norm() const { return c_.norm(); }
    void normalize() { c_.normalize(); }
    Quaternion normalized() const { Quaternion q(*this); q.normalize(); return q; }
    Quaternion conjugate() const { return Quaternion(w(), -x(), -y(), -z()); }
    Quaternion inverse() const { const S n2 = squaredNorm(); if (n2 > S(0)) { Quaternion q = conjugate(); q.c_ /= n2; return q; } Quaternion q; q.c_.setZero(); return q; }
    // Eigen/src/Geometry/Quaternion.h: quaternionbase_assign_impl<Other, 3, 3>
    template <class Dv> Quaternion &operator=(const MatrixBase<Dv> &mat)
    {
        if (mat.rows() == 4 && mat.cols() == 1) { for (int i = 0; i < 4; i++) c_[i] = mat.coeff(i, 0); return *this; }
        S t = mat.trace();
        if (t > S(0)) {
            t = std::sqrt(t + S(1.0));
            w() = S(0.5) * t;
            t = S(0.5) / t;
            x() = (mat.coeff(2, 1) - mat.coeff(1, 2)) * t; y() = (mat.coeff(0, 2) - mat.coeff(2, 0)) * t; z() = (mat.coeff(1, 0) - mat.coeff(0, 1)) * t;
        } else {
            Index i = 0;
            if (mat.coeff(1, 1) > mat.coeff(0, 0)) i = 1;
            if (mat.coeff(2, 2) > mat.coeff(i, i)) i = 2;
            const Index j = (i + 1) % 3, k = (j + 1) % 3;
            t = std::sqrt(mat.coeff(i, i) - mat.coeff(j, j) - mat.coeff(k, k) + S(1.0));
            c_[i] = S(0.5) * t;
            t = S(0.5) / t;
            w() = (mat.coeff(k, j) - mat.coeff(j, k)) * t;
            c_[j] = (mat.coeff(j, i) + mat.coeff(i, j)) * t;
            c_[k] = (mat.coeff(k, i) + mat.coeff(i, k)) * t;
        }
        return *this;
    }
    Matrix3 toRotationMatrix() const
    {
        Matrix3 res;
        const S tx = S(2) * x(), ty = S(2) * y(), tz = S(2) * z();
        const S twx = tx * w(), twy = ty * w(), twz = tz * w();
        const S txx = tx * x(), txy = ty * x(), txz = tz * x();
        const S tyy = ty * y(), tyz = tz * y(), tzz = tz * z();
        res.coeffRef(0, 0) = S(1) - (tyy + tzz); res.coeffRef(0, 1) = txy - twz; res.coeffRef(0, 2) = txz + twy;
        res.coeffRef(1, 0) = txy + twz; res.coeffRef(1, 1) = S(1) - (txx + tzz); res.coeffRef(1, 2) = tyz - twx;
        res.coeffRef(2, 0) = txz - twy; res.coeffRef(2, 1) = tyz + twx; res.coeffRef(2, 2) = S(1) - (txx + tyy);
        return res;
    }
    Quaternion operator*(const Quaternion &b) const
    {
        const Quaternion &a = *this;
        return Quaternion(a.w() * b.w() - a.x() * b.x() - a.y() * b.y() - a.z() * b.z(), a.w() * b.x() + a.x() * b.w() + a.y() * b.z() - a.z() * b.y(),
                          a.w() * b.y() + a.y() * b.w() + a.z() * b.x() - a.x() * b.z(), a.w() * b.z() + a.z() * b.w() + a.x() * b.y() - a.y() * b.x());
    }
    Quaternion &operator*=(const Quaternion &b) { *this = *this * b; return *this; }
    // _transformVector: v + w * uv + vec x uv with uv = 2 (vec x v)
    template <class Dv> Vector3 _transformVector(const MatrixBase<Dv> &v) const
    {
        Vector3 uv = vec().cross(v);
        uv += uv;
        const Vector3 c2 = vec().cross(uv);
        return Vector3(v.coeff(0) + w() * uv[0] + c2[0], v.coeff(1) + w() * uv[1] + c2[1], v.coeff(2) + w() * uv[2] + c2[2]);
    }
    template <class Dv> Vector3 operator*(const MatrixBase<Dv> &v) const { return _transformVector(v); }
    template <class NewS> Quaternion<NewS> cast() const { return Quaternion<NewS>(NewS(w()), NewS(x()), NewS(y()), NewS(z())); }
};
typedef Quaternion<double> Quaterniond; typedef Quaternion<float> Quaternionf;

template <class S> class AngleAxis {
    Matrix<S, 3, 1> axis_; S angle_;
  public:
    AngleAxis() : angle_(0) {}
    template <class Dv> AngleAxis(const S &angle, const MatrixBase<Dv> &axis) : axis_(axis), angle_(angle) {}
    S angle() const { return angle_; }
    const Matrix<S, 3, 1> &axis() const { return axis_; }
    Matrix<S, 3, 3> toRotationMatrix() const { return Quaternion<S>(*this).toRotationMatrix(); }
};
typedef AngleAxis<double> AngleAxisd;
template <class S, int O> Quaternion<S, O>::Quaternion(const AngleAxis<S> &aa) { const S ha = S(0.5) * aa.angle(); w() = std::cos(ha); const S s = std::sin(ha); x() = s * aa.axis()[0]; y() = s * aa.axis()[1]; z() = s * aa.axis()[2]; }

template <class S, int Dim, int Mode, int Opt = 0> class Transform {
    Matrix<S, Dim + 1, Dim + 1> m_;
  public:
    Transform() { m_.setIdentity(); }
    Transform(const Quaternion<S> &q) { m_.setIdentity(); m_.template block<Dim, Dim>(0, 0) = q.toRotationMatrix(); }
    template <class Dv> explicit Transform(const MatrixBase<Dv> &m) { m_.setIdentity(); if (m.rows() == Dim) m_.template block<Dim, Dim>(0, 0) = m; else m_ = m; }
    static Transform Identity() { return Transform(); }
    Block<Matrix<S, Dim + 1, Dim + 1>, Dim, 1> translation() { return m_.template block<Dim, 1>(0, Dim); }
    const Block<const Matrix<S, Dim + 1, Dim + 1>, Dim, 1> translation() const { return m_.template block<Dim, 1>(0, Dim); }
    Block<Matrix<S, Dim + 1, Dim + 1>, Dim, Dim> linear() { return m_.template block<Dim, Dim>(0, 0); }
    const Block<const Matrix<S, Dim + 1, Dim + 1>, Dim, Dim> linear() const { return m_.template block<Dim, Dim>(0, 0); }
    Matrix<S, Dim, Dim> rotation() const { return Matrix<S, Dim, Dim>(linear()); }
    const Matrix<S, Dim + 1, Dim + 1> &matrix() const { return m_; }
    Matrix<S, Dim + 1, Dim + 1> &matrix() { return m_; }
    Transform operator*(const Transform &o) const { Transform r; r.m_ = m_ * o.m_; return r; }
    template <class Dv> Matrix<S, Dim, 1> operator*(const MatrixBase<Dv> &v) const { return Matrix<S, Dim, 1>(linear() * v + translation()); }
    Transform inverse() const { Transform r; r.m_ = m_.inverse(); return r; }
};
typedef Transform<double, 3, Isometry> Isometry3d; typedef Transform<double, 2, Isometry> Isometry2d;
typedef Transform<double, 3, Affine> Affine3d; typedef Transform<double, 2, Affine> Affine2d;

// ---- Sparse: column-compressed matrix, triplets, permutation, simplicial LDL^T (upper triangle in, as g2o's LinearSolverEigen feeds it)
template <class S> class Triplet {
    Index r_, c_; S v_;
  public:
    Triplet() : r_(0), c_(0), v_(0) {}
    Triplet(Index r, Index c, const S &v = S(0)) : r_(r), c_(c), v_(v) {}
    Index row() const { return r_; } Index col() const { return c_; } const S &value() const { return v_; }
};

template <int SR = Dynamic, int SC = Dynamic, class I = int> class PermutationMatrix {
    Matrix<int, Dynamic, 1> idx_;
  public:
    PermutationMatrix() {}
    explicit PermutationMatrix(Index n) { resize(n); }
    void resize(Index n) { idx_.resize(n); }
    Index size() const { return idx_.size(); }
    Matrix<int, Dynamic, 1> &indices() { return idx_; }
    const Matrix<int, Dynamic, 1> &indices() const { return idx_; }
    void setIdentity(Index n) { resize(n); for (Index i = 0; i < n; i++) idx_(i) = (int)i; }
    PermutationMatrix inverse() const { PermutationMatrix r(size()); for (Index i = 0; i < size(); i++) r.idx_(idx_(i)) = (int)i; return r; }
};

template <class SM, int UpLo> class SparseSelfAdjointView;
template <class S, int Opt = ColMajor, class I = int> class SparseMatrix {
    Index r_, c_;
    std::vector<I> outer_, inner_;
    std::vector<S> val_;
  public:
    typedef S Scalar; typedef I StorageIndex; typedef I Index_;
    SparseMatrix() : r_(0), c_(0), outer_(1, 0) {}
    SparseMatrix(Index r, Index c) : r_(r), c_(c), outer_((size_t)c + 1, 0) {}
    void resize(Index r, Index c) { r_ = r; c_ = c; outer_.assign((size_t)c + 1, 0); inner_.clear(); val_.clear(); }
    Index rows() const { return r_; } Index cols() const { return c_; }
    Index nonZeros() const { return (Index)val_.size(); }
    S *valuePtr() { return val_.data(); } const S *valuePtr() const { return val_.data(); }
    I *innerIndexPtr() { return inner_.data(); } const I *innerIndexPtr() const { return inner_.data(); }
    I *outerIndexPtr() { return outer_.data(); } const I *outerIndexPtr() const { return outer_.data(); }
    void makeCompressed() {}
    // duplicates are summed; entries sorted by (column, row) -- the order SparseBlockMatrix::fillCCS relies on
    template <class It> void setFromTriplets(It b, It e)
    {
        std::vector<std::pair<std::pair<Index, Index>, S> > t;
        for (It it = b; it != e; ++it) t.push_back(std::make_pair(std::make_pair((Index)it->col(), (Index)it->row()), (S)it->value()));
        std::stable_sort(t.begin(), t.end(), [](const std::pair<std::pair<Index, Index>, S> &x, const std::pair<std::pair<Index, Index>, S> &y) { return x.first < y.first; });
        outer_.assign((size_t)c_ + 1, 0); inner_.clear(); val_.clear();
        for (size_t q = 0; q < t.size(); q++) {
            if (q > 0 && t[q].first == t[q - 1].first) { val_.back() += t[q].second; continue; }
            inner_.push_back((I)t[q].first.second); val_.push_back(t[q].second); outer_[(size_t)t[q].first.first + 1]++;
        }
        for (Index j = 0; j < c_; j++) outer_[(size_t)j + 1] += outer_[(size_t)j];
    }
    template <int UpLo> SparseSelfAdjointView<SparseMatrix, UpLo> selfadjointView() { return SparseSelfAdjointView<SparseMatrix, UpLo>(*this); }
    template <int UpLo> SparseSelfAdjointView<const SparseMatrix, UpLo> selfadjointView() const { return SparseSelfAdjointView<const SparseMatrix, UpLo>(*this); }
    template <class SMx, int U> SparseMatrix &operator=(const SparseSelfAdjointView<SMx, U> &v);
};
template <class SM, int UpLo> class SparseSelfAdjointView {
  public:
    SM &m_; const PermutationMatrix<> *p_;
    explicit SparseSelfAdjointView(SM &m, const PermutationMatrix<> *p = nullptr) : m_(m), p_(p) {}
    SparseSelfAdjointView twistedBy(const PermutationMatrix<> &p) const { return SparseSelfAdjointView(m_, &p); }
    // upper-triangle to upper-triangle copy under a symmetric permutation (new index = pinv[old index], as Eigen's twistedBy(P) with P = perm.inverse())
    SparseSelfAdjointView(const SparseSelfAdjointView &) = default;
    SparseSelfAdjointView &operator=(const SparseSelfAdjointView &src) { return this->template assign_from<SM, UpLo>(src); }
    template <class SMx, int U> SparseSelfAdjointView &operator=(const SparseSelfAdjointView<SMx, U> &src) { return this->template assign_from<SMx, U>(src); }
    template <class SMx, int U> SparseSelfAdjointView &assign_from(const SparseSelfAdjointView<SMx, U> &src)
    {
        typedef typename std::remove_const<SM>::type M;
        std::vector<Triplet<typename M::Scalar> > t;
        const Index n = src.m_.cols();
        std::vector<int> pinv((size_t)n);
        for (Index i = 0; i < n; i++) pinv[(size_t)i] = (int)i;
        if (src.p_) { PermutationMatrix<> inv = src.p_->inverse(); for (Index i = 0; i < n; i++) pinv[(size_t)i] = inv.indices()(i); }
        for (Index j = 0; j < n; j++)
            for (int q = src.m_.outerIndexPtr()[j]; q < src.m_.outerIndexPtr()[j + 1]; q++) {
                const Index i = src.m_.innerIndexPtr()[q];
                if (i > j) continue;
                const Index a = pinv[(size_t)i], b = pinv[(size_t)j];
                t.push_back(Triplet<typename M::Scalar>(std::min(a, b), std::max(a, b), src.m_.valuePtr()[q]));
            }
        const_cast<M &>(m_).resize(n, n);
        const_cast<M &>(m_).setFromTriplets(t.begin(), t.end());
        return *this;
    }
};
template <class S, int Opt, class I> template <class SMx, int U> SparseMatrix<S, Opt, I> &SparseMatrix<S, Opt, I>::operator=(const SparseSelfAdjointView<SMx, U> &v)
{
    // full symmetric matrix from the stored triangle
    std::vector<Triplet<S> > t;
    for (Index j = 0; j < v.m_.cols(); j++)
        for (int q = v.m_.outerIndexPtr()[j]; q < v.m_.outerIndexPtr()[j + 1]; q++) {
            const Index i = v.m_.innerIndexPtr()[q];
            t.push_back(Triplet<S>(i, j, v.m_.valuePtr()[q]));
            if (i != j) t.push_back(Triplet<S>(j, i, v.m_.valuePtr()[q]));
        }
    resize(v.m_.rows(), v.m_.cols());
    setFromTriplets(t.begin(), t.end());
    return *this;
}

namespace internal {
// fill-reducing ordering stand-in: reverse Cuthill-McKee on the symmetric pattern (Eigen uses AMD; any ordering gives the same solution up to
// rounding).  perm.indices()(new) = old.
template <class SM> void minimum_degree_ordering(SM &C, PermutationMatrix<> &perm)
{
    const Index n = C.cols();
    std::vector<std::vector<int> > adj((size_t)n);
    for (Index j = 0; j < n; j++) for (int q = C.outerIndexPtr()[j]; q < C.outerIndexPtr()[j + 1]; q++) { const int i = C.innerIndexPtr()[q]; if (i != j) { adj[(size_t)i].push_back((int)j); adj[(size_t)j].push_back(i); } }
    for (auto &a : adj) { std::sort(a.begin(), a.end()); a.erase(std::unique(a.begin(), a.end()), a.end()); }
    std::vector<int> order; order.reserve((size_t)n);
    std::vector<char> seen((size_t)n, 0);
    for (Index s0 = 0; s0 < n; s0++) {
        if (seen[(size_t)s0]) continue;
        // start from a minimum-degree vertex of this component
        std::vector<int> comp(1, (int)s0), stack(1, (int)s0); seen[(size_t)s0] = 2;
        while (!stack.empty()) { const int u = stack.back(); stack.pop_back(); for (int v : adj[(size_t)u]) if (!seen[(size_t)v]) { seen[(size_t)v] = 2; stack.push_back(v); comp.push_back(v); } }
        int start = comp[0];
        for (int v : comp) { if (adj[(size_t)v].size() < adj[(size_t)start].size()) start = v; seen[(size_t)v] = 0; }
        size_t head = order.size();
        order.push_back(start); seen[(size_t)start] = 1;
        while (head < order.size()) {
            const int u = order[head++];
            std::vector<int> nb;
            for (int v : adj[(size_t)u]) if (!seen[(size_t)v]) { seen[(size_t)v] = 1; nb.push_back(v); }
            std::sort(nb.begin(), nb.end(), [&](int a, int b) { return adj[(size_t)a].size() != adj[(size_t)b].size() ? adj[(size_t)a].size() < adj[(size_t)b].size() : a < b; });
            order.insert(order.end(), nb.begin(), nb.end());
        }
    }
    std::reverse(order.begin(), order.end());
    // test knob: EIGENSHIM_ORDERING=natural | reverse replaces the fill-reducing ordering (used to measure how far the REFERENCE's own results move with
    // the elimination order, i.e. with the Eigen version it happens to be built against)
    if (const char *o = std::getenv("EIGENSHIM_ORDERING")) {
        if (!std::strcmp(o, "natural")) for (Index i = 0; i < n; i++) order[(size_t)i] = (int)i;
        else if (!std::strcmp(o, "reverse")) for (Index i = 0; i < n; i++) order[(size_t)i] = (int)(n - 1 - i);
    }
    perm.resize(n);
    for (Index i = 0; i < n; i++) perm.indices()(i) = order[(size_t)i];
}
}  // namespace internal

// Up-looking sparse LDL^T (T. Davis' LDL: elimination tree + row patterns), no numerical pivoting -- like Eigen::SimplicialLDLT.
template <class SM, int UpLo_ = Lower> class SimplicialLDLT {
  public:
    enum { UpLo = UpLo_ };
    typedef SM CholMatrixType;
    typedef typename SM::Scalar S;
    struct LView { const SimplicialLDLT *s; const LView &nestedExpression() const { return *this; } Index nonZeros() const { return (Index)s->Li_.size(); } };
  protected:
    PermutationMatrix<> m_P, m_Pinv;          // m_P.indices()(new) = old
    ComputationInfo info_;
    Index n_;
    std::vector<int> parent_, Lp_, Li_, Cp_, Ci_;
    std::vector<S> Lx_, D_, Cx_;
    // C = upper triangle of the symmetrically permuted matrix, column-compressed
    void permute_upper(const SM &a)
    {
        const Index n = a.cols();
        std::vector<int> pinv((size_t)n);
        for (Index i = 0; i < n; i++) pinv[(size_t)m_P.indices()(i)] = (int)i;
        std::vector<int> cnt((size_t)n + 1, 0);
        for (Index j = 0; j < n; j++) for (int q = a.outerIndexPtr()[j]; q < a.outerIndexPtr()[j + 1]; q++) { const int i = a.innerIndexPtr()[q]; if (i > j) continue; cnt[(size_t)std::max(pinv[(size_t)i], pinv[(size_t)j]) + 1]++; }
        Cp_.assign((size_t)n + 1, 0);
        for (Index j = 0; j < n; j++) Cp_[(size_t)j + 1] = Cp_[(size_t)j] + cnt[(size_t)j + 1];
        Ci_.assign((size_t)Cp_[(size_t)n], 0); Cx_.assign((size_t)Cp_[(size_t)n], S(0));
        std::vector<int> fill(Cp_.begin(), Cp_.end() - 1);
        for (Index j = 0; j < n; j++) for (int q = a.outerIndexPtr()[j]; q < a.outerIndexPtr()[j + 1]; q++) {
            const int i = a.innerIndexPtr()[q]; if (i > j) continue;
            const int pi = pinv[(size_t)i], pj = pinv[(size_t)j], col = std::max(pi, pj), row = std::min(pi, pj);
            Ci_[(size_t)fill[(size_t)col]] = row; Cx_[(size_t)fill[(size_t)col]++] = a.valuePtr()[q];
        }
    }
    void symbolic()
    {
        const Index n = n_;
        parent_.assign((size_t)n, -1);
        std::vector<int> flag((size_t)n), lnz((size_t)n, 0);
        for (Index k = 0; k < n; k++) {
            flag[(size_t)k] = (int)k;
            for (int p = Cp_[(size_t)k]; p < Cp_[(size_t)k + 1]; p++) {
                int i = Ci_[(size_t)p];
                if (i >= k) continue;
                for (; flag[(size_t)i] != k; i = parent_[(size_t)i]) { if (parent_[(size_t)i] == -1) parent_[(size_t)i] = (int)k; lnz[(size_t)i]++; flag[(size_t)i] = (int)k; }
            }
        }
        Lp_.assign((size_t)n + 1, 0);
        for (Index k = 0; k < n; k++) Lp_[(size_t)k + 1] = Lp_[(size_t)k] + lnz[(size_t)k];
        Li_.assign((size_t)Lp_[(size_t)n], 0); Lx_.assign((size_t)Lp_[(size_t)n], S(0)); D_.assign((size_t)n, S(0));
    }
    void analyzePattern_preordered(const SM &ap, bool)
    {
        // ap is already permuted: keep its pattern as C with the identity on top of m_P
        n_ = ap.cols();
        Cp_.assign(ap.outerIndexPtr(), ap.outerIndexPtr() + n_ + 1); Ci_.assign(ap.innerIndexPtr(), ap.innerIndexPtr() + ap.nonZeros()); Cx_.assign((size_t)ap.nonZeros(), S(0));
        symbolic();
    }
  public:
    SimplicialLDLT() : info_(Success), n_(0) {}
    void analyzePattern(const SM &a)
    {
        n_ = a.cols();
        SM full; full = a.template selfadjointView<UpLo>();
        internal::minimum_degree_ordering(full, m_P);
        m_Pinv = m_P.inverse();
        permute_upper(a);
        symbolic();
    }
    void factorize(const SM &a)
    {
        permute_upper(a);
        const Index n = n_;
        std::vector<S> Y((size_t)n, S(0));
        std::vector<int> pattern((size_t)n), flag((size_t)n), lnz((size_t)n, 0);
        info_ = Success;
        for (Index k = 0; k < n; k++) {
            Y[(size_t)k] = S(0);
            int top = (int)n;
            flag[(size_t)k] = (int)k;
            for (int p = Cp_[(size_t)k]; p < Cp_[(size_t)k + 1]; p++) {
                int i = Ci_[(size_t)p];
                if (i > k) continue;
                Y[(size_t)i] += Cx_[(size_t)p];
                int len = 0;
                for (; flag[(size_t)i] != k; i = parent_[(size_t)i]) { pattern[(size_t)len++] = i; flag[(size_t)i] = (int)k; }
                while (len > 0) pattern[(size_t)--top] = pattern[(size_t)--len];
            }
            D_[(size_t)k] = Y[(size_t)k]; Y[(size_t)k] = S(0);
            for (; top < n; top++) {
                const int i = pattern[(size_t)top];
                const S yi = Y[(size_t)i]; Y[(size_t)i] = S(0);
                const int p2 = Lp_[(size_t)i] + lnz[(size_t)i];
                for (int p = Lp_[(size_t)i]; p < p2; p++) Y[(size_t)Li_[(size_t)p]] -= Lx_[(size_t)p] * yi;
                const S lki = yi / D_[(size_t)i];
                D_[(size_t)k] -= lki * yi;
                Li_[(size_t)p2] = (int)k; Lx_[(size_t)p2] = lki; lnz[(size_t)i]++;
            }
            if (D_[(size_t)k] == S(0)) { info_ = NumericalIssue; return; }
        }
    }
    void compute(const SM &a) { analyzePattern(a); factorize(a); }
    ComputationInfo info() const { return info_; }
    LView matrixL() const { LView v; v.s = this; return v; }
    template <class B> VectorXd solve(const MatrixBase<B> &b) const
    {
        const Index n = n_;
        std::vector<S> x((size_t)n);
        for (Index i = 0; i < n; i++) x[(size_t)i] = b.coeff(m_P.indices()(i));
        for (Index j = 0; j < n; j++) for (int p = Lp_[(size_t)j]; p < Lp_[(size_t)j + 1]; p++) x[(size_t)Li_[(size_t)p]] -= Lx_[(size_t)p] * x[(size_t)j];
        for (Index j = 0; j < n; j++) x[(size_t)j] /= D_[(size_t)j];
        for (Index j = n - 1; j >= 0; j--) for (int p = Lp_[(size_t)j]; p < Lp_[(size_t)j + 1]; p++) x[(size_t)j] -= Lx_[(size_t)p] * x[(size_t)Li_[(size_t)p]];
        VectorXd r(n);
        for (Index i = 0; i < n; i++) r(m_P.indices()(i)) = x[(size_t)i];
        return r;
    }
};

}  // namespace Eigen
