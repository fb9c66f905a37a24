// smm_b200.hpp -- drop-in C++17 surface (namespace SMM) over the B200 C ABI (smm_b200.h).
//
// A user of vasil-pashov/sparse_matrix_math includes <sparse_matrix_math.h> (this repo ships a forwarding header of
// that name) and links libsmm_b200.so.  Containers, iterators and file I/O are host code written from scratch with
// the reference's observable behaviour; every function on the Krylov hot path forwards to the C ABI and runs on the
// GPU for T = float:
//     CSRMatrix<float>::rMult / rMultAdd / rMultSub                       -> smm_spmv            (ref H:1458-1515)
//     Vector<float>::operator* , secondNorm[Squared]                      -> smm_dot             (ref H:287-328)
//     CSRMatrix<float>::getPreconditioner<SYMMETRIC_GAUS_SEIDEL>().apply  -> smm_precond_apply   (ref H:1643-1713)
//     ConjugateGradient / BiCGSymmetric / ConjugateGradientSquared / BiCGStab -> smm_solve_*     (ref H:2021-2398)
// ("ref H:n" = line n of the reference's include/sparse_matrix_math.h.)  There is no CPU implementation of these
// behind the header: for any other scalar type they do not compile (static_assert), and at run time a missing GPU
// or library aborts with the C ABI's error text.
//
// Additive (not in the reference): CSRMatrix::init(rows, cols, start, positions, values) for direct CSR ingest,
// SMM::SolveInfo / SMM::b200::lastSolveInfo() for iteration counts, SMM::b200::options() for the reduction and
// driver modes, the README spellings ConjugateGradientSqared and SYMMETRIC_GAUSS_SEIDEL.
// Multi-GPU: SMM::b200::devices() = N (default 1) row-partitions the matrix over the first N GPUs of the box inside THIS
// process (smm_group_*, one host thread per GPU, P2P halo exchange and reductions between the kernels): rMult*, ConjugateGradient,
// BiCGSymmetric, ConjugateGradientSquared and the unpreconditioned BiCGStab then run on N GPUs with the same signatures.
// Preconditioned solves stay on one GPU (triangular sweeps do not shard exactly).
#pragma once

#include <algorithm>
#include <cassert>
#include <cctype>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <initializer_list>
#include <iomanip>
#include <iterator>
#include <limits>
#include <map>
#include <memory>
#include <string>
#include <type_traits>
#include <unordered_map>
#include <utility>
#include <vector>

#include "smm_b200.h"

#define SMM_MAJOR_VERSION 0
#define SMM_MINOR_VERSION 2
#define SMM_PATCH_VERSION 0
#define SMM_B200 1

namespace SMM {

// ---------------------------------------------------------------------------------------------------------------
// B200 glue
// ---------------------------------------------------------------------------------------------------------------
struct SolveInfo {
    int iterations = 0;
    float residual = 0.0f;       // ||r||^2 for CG / BiCGSymmetric / CGS, ||r||_2 for BiCGStab (what the solver compares)
    int precondError = 0;
    double secondsSolve = 0.0;   // device time, vectors resident
    double secondsTotal = 0.0;   // including host<->device copies
    long long kernelLaunches = 0;
};

namespace b200 {
inline smm_solve_options& options() {
    static smm_solve_options o = {SMM_REDUCE_FAST, SMM_DRIVER_AUTO, 0, 0, nullptr, {0, 0, 0, 0}};
    return o;
}
// number of GPUs the unpreconditioned hot path uses (one process, row blocks; see the header comment)
inline int& devices() {
    static int n = 1;
    return n;
}
inline SolveInfo& lastSolveInfo() {
    static thread_local SolveInfo info;
    return info;
}
[[noreturn]] inline void die(const char* what, int rc) {
    std::fprintf(stderr, "sparse_matrix_math (B200): %s failed with code %d: %s\n", what, rc, smm_last_error());
    std::abort();
}
inline void check(int rc, const char* what) {
    if (rc != SMM_OK) die(what, rc);
}
template <typename T>
constexpr void requireFloat() {
    static_assert(std::is_same_v<T, float>, "the B200 Krylov path is implemented for T = float only (no CPU fallback)");
}
inline void record(const smm_solve_info& i) {
    SolveInfo& o = lastSolveInfo();
    o.iterations = i.iterations;
    o.residual = i.residual;
    o.precondError = i.precond_error;
    o.secondsSolve = i.seconds_solve;
    o.secondsTotal = i.seconds_total;
    o.kernelLaunches = i.kernel_launches;
}
}  // namespace b200

template <typename T>
using do_not_deduce = std::common_type_t<T>;

// ---------------------------------------------------------------------------------------------------------------
// Vector<T>: malloc-backed, move-only dense vector (ref H:42-381)
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
class Vector {
public:
    using Iterator = T*;
    using ConstIterator = const T*;

    Vector() noexcept : data(nullptr), size(0) {}
    explicit Vector(const int n) noexcept : data(static_cast<T*>(std::malloc(sizeof(T) * (n > 0 ? n : 0)))), size(n) {}
    Vector(const int n, const do_not_deduce<T> val) noexcept : data(nullptr), size(n) { allocateFilled(val); }
    Vector(const std::initializer_list<T>& l) noexcept
        : data(static_cast<T*>(std::malloc(sizeof(T) * l.size()))), size(static_cast<int>(l.size())) {
        std::copy(l.begin(), l.end(), data);
    }
    Vector(Vector&& o) noexcept : data(o.data), size(o.size) { o.data = nullptr; o.size = 0; }
    Vector& operator=(Vector&& o) noexcept {
        if (this != &o) {
            std::free(data);
            data = o.data; size = o.size;
            o.data = nullptr; o.size = 0;
        }
        return *this;
    }
    Vector(const Vector&) = delete;
    Vector& operator=(const Vector&) = delete;
    ~Vector() { deinit(); }

    void init(const int n) {
        if (size < n) {
            data = static_cast<T*>(std::realloc(data, sizeof(T) * n));
            assert(data != nullptr);
        }
        size = n;
    }
    void init(const int n, const T val) { init(n); fill(val); }
    void deinit() noexcept { std::free(data); data = nullptr; size = 0; }
    int getSize() const { return size; }
    operator T*() { return data; }
    const T& operator[](const int i) const { assert(i >= 0 && i < size); return data[i]; }
    T& operator[](const int i) { assert(i >= 0 && i < size); return data[i]; }

    Vector& operator+=(const Vector& o) {
        assert(o.size == size);
        for (int i = 0; i < size; ++i) data[i] += o.data[i];
        return *this;
    }
    Vector& operator-=(const Vector& o) {
        assert(o.size == size);
        for (int i = 0; i < size; ++i) data[i] -= o.data[i];
        return *this;
    }

    // dot product and norms: reductions of the hot path -> GPU (float).  The reference's summation orders are kept:
    // operator* = the SMM_MULTITHREADING build's deterministic reduce tree, the norms = left to right.
    const T operator*(const Vector& o) const {
        b200::requireFloat<T>();
        assert(o.size == size);
        float r = 0.0f;
        b200::check(smm_dot(size, data, o.data, SMM_REDUCE_REFERENCE_TREE, &r), "smm_dot");
        return r;
    }
    T secondNormSquared() const {
        b200::requireFloat<T>();
        float r = 0.0f;
        b200::check(smm_dot(size, data, data, SMM_REDUCE_REFERENCE_SERIAL, &r), "smm_dot");
        return r;
    }
    T secondNorm() const { return std::sqrt(secondNormSquared()); }

    Iterator begin() noexcept { return data; }
    Iterator end() noexcept { return data + size; }
    ConstIterator begin() const noexcept { return data; }
    ConstIterator end() const noexcept { return data + size; }
    ConstIterator cbegin() const noexcept { return data; }
    ConstIterator cend() const noexcept { return data + size; }

    void fill(const T value) {
        if (value == T(0)) std::memset(data, 0, sizeof(T) * size);
        else std::fill_n(data, size, value);
    }
    void swap(Vector& o) { std::swap(data, o.data); std::swap(size, o.size); }

private:
    void allocateFilled(const T val) {
        if (val == T(0)) {
            data = static_cast<T*>(std::calloc(size > 0 ? size : 0, sizeof(T)));
        } else {
            data = static_cast<T*>(std::malloc(static_cast<std::size_t>(size > 0 ? size : 0) * sizeof(T)));
            if (data) std::fill_n(data, size, val);
        }
    }
    T* data;
    int size;
};

// ---------------------------------------------------------------------------------------------------------------
// Triplet (coordinate) builders (ref H:383-684): key = (row << 32) | col, duplicates summed in call order,
// explicit zeros kept.  Host structures; the ordered one feeds CSRMatrix.
// ---------------------------------------------------------------------------------------------------------------
template <typename Container, typename T>
class TripletMatrixConstIterator;
template <typename Container, typename T>
class _TripletMatrixCommon;

template <typename Container, typename T>
class TripletEl {
    template <typename, typename> friend class TripletMatrixConstIterator;
public:
    TripletEl(const TripletEl&) noexcept = default;
    TripletEl& operator=(const TripletEl&) = default;
    bool operator==(const TripletEl& o) const { return it == o.it; }
    int getRow() const noexcept { return static_cast<int>(it->first >> 32); }
    int getCol() const noexcept { return static_cast<int>(it->first & 0xFFFFFFFFu); }
    const T getValue() const noexcept { return it->second; }
    friend void swap(TripletEl& a, TripletEl& b) noexcept { using std::swap; swap(a.it, b.it); }
private:
    explicit TripletEl(typename Container::const_iterator i) : it(i) {}
    typename Container::const_iterator it;
};

template <typename Container, typename T>
class TripletMatrixConstIterator {
    template <typename, typename> friend class _TripletMatrixCommon;
public:
    using iterator_category = std::forward_iterator_tag;
    using value_type = TripletEl<Container, T>;
    using difference_type = std::ptrdiff_t;
    using pointer = const value_type*;
    using reference = const value_type&;
    bool operator==(const TripletMatrixConstIterator& o) const noexcept { return el == o.el; }
    bool operator!=(const TripletMatrixConstIterator& o) const noexcept { return !(el == o.el); }
    reference operator*() const { return el; }
    pointer operator->() const { return &el; }
    TripletMatrixConstIterator& operator++() noexcept { ++el.it; return *this; }
    TripletMatrixConstIterator operator++(int) noexcept { TripletMatrixConstIterator t = *this; ++el.it; return t; }
private:
    explicit TripletMatrixConstIterator(typename Container::const_iterator i) : el(i) {}
    value_type el;
};

template <typename Container, typename T>
class _TripletMatrixCommon {
public:
    using value_type = T;
    using ConstIterator = TripletMatrixConstIterator<Container, T>;
    _TripletMatrixCommon() : rows(0), cols(0) {}
    _TripletMatrixCommon(int rowCount, int colCount) noexcept : rows(rowCount), cols(colCount) {}
    _TripletMatrixCommon(int rowCount, int colCount, int /*numTriplets*/) noexcept : rows(rowCount), cols(colCount) {}
    _TripletMatrixCommon(_TripletMatrixCommon&&) = default;
    _TripletMatrixCommon& operator=(_TripletMatrixCommon&&) = default;
    _TripletMatrixCommon(const _TripletMatrixCommon&) = delete;
    _TripletMatrixCommon& operator=(const _TripletMatrixCommon&) = delete;

    void init(int rowCount, int colCount, int /*numTriplets*/) {
        assert(entries.empty() && rows == 0 && cols == 0);
        rows = rowCount; cols = colCount;
    }
    void deinit() { rows = 0; cols = 0; entries.clear(); }
    void addEntry(int row, int col, T value) {
        assert(row >= 0 && row < rows && col >= 0 && col < cols);
        auto ins = entries.emplace(key(row, col), value);
        if (!ins.second) ins.first->second += value;
    }
    bool updateEntry(const int row, const int col, const T newValue) {
        auto it = entries.find(key(row, col));
        if (it == entries.end()) return false;
        it->second = newValue;
        return true;
    }
    T getValue(const int row, const int col) const {
        auto it = entries.find(key(row, col));
        return it == entries.end() ? T(0) : it->second;
    }
    ConstIterator begin() const noexcept { return ConstIterator(entries.cbegin()); }
    ConstIterator end() const noexcept { return ConstIterator(entries.cend()); }
    int getNonZeroCount() const noexcept { return static_cast<int>(entries.size()); }
    int getDenseRowCount() const noexcept { return rows; }
    int getDenseColCount() const noexcept { return cols; }
    _TripletMatrixCommon& operator*=(const T scalar) noexcept {
        for (auto& kv : entries) kv.second *= scalar;
        return *this;
    }
private:
    static std::uint64_t key(int row, int col) {
        static_assert(sizeof(int) == 4, "32-bit int expected");
        return (static_cast<std::uint64_t>(static_cast<std::uint32_t>(row)) << 32) | static_cast<std::uint32_t>(col);
    }
    Container entries;
    int rows, cols;
};

template <typename T>
using TripletMatrix = _TripletMatrixCommon<std::map<std::uint64_t, T>, T>;
template <typename T>
using UnorderedTripletMatrix = _TripletMatrixCommon<std::unordered_map<std::uint64_t, T>, T>;

// ---------------------------------------------------------------------------------------------------------------
// CSR iterators (ref H:686-1000).  MatrixPtrT is `CSRMatrix<T>*` or `const CSRMatrix<T>*`.
// ---------------------------------------------------------------------------------------------------------------
template <typename P>
struct is_ptr_to_const : std::bool_constant<std::is_pointer_v<P> && std::is_const_v<std::remove_pointer_t<P>>> {};
template <typename P>
inline constexpr bool is_ptr_to_const_t = is_ptr_to_const<P>::value;
template <typename P>
using make_ptr_to_const_t = const std::remove_pointer_t<P>*;

template <typename MatrixPtrT>
class _CSRIteratorBase {
public:
    static_assert(std::is_pointer_v<MatrixPtrT>, "matrix pointer type expected");
    using el_value_type = typename std::remove_pointer_t<MatrixPtrT>::value_type;
    template <typename> friend class _CSRIteratorBase;

    class CSRElement {
    public:
        template <typename> friend class _CSRIteratorBase;
        CSRElement(MatrixPtrT matrix, int row, int index) noexcept : m(matrix), row(row), index(index) {}
        CSRElement(const CSRElement&) = default;
        CSRElement& operator=(const CSRElement&) = default;
        const el_value_type getValue() const noexcept { return m->values[index]; }
        int getRow() const noexcept { return row; }
        int getCol() const noexcept { return m->positions[index]; }
        void setValue(el_value_type v) noexcept { m->values[index] = v; m->touchValues(); }
        bool operator==(const CSRElement& o) const { return m == o.m && row == o.row && index == o.index; }
    protected:
        MatrixPtrT m;
        int row;     // index into start
        int index;   // index into positions / values
    };

    using iterator_category = std::forward_iterator_tag;
    using value_type = CSRElement;
    using difference_type = std::ptrdiff_t;
    using pointer = std::conditional_t<is_ptr_to_const_t<MatrixPtrT>, const CSRElement*, CSRElement*>;
    using reference = std::conditional_t<is_ptr_to_const_t<MatrixPtrT>, const CSRElement&, CSRElement&>;

    _CSRIteratorBase(MatrixPtrT m, int row, int index) noexcept : cur(m, row, index) {}
    template <typename Other, typename = std::enable_if_t<!std::is_same_v<Other, MatrixPtrT> && (!is_ptr_to_const_t<Other> || is_ptr_to_const_t<MatrixPtrT>)>>
    _CSRIteratorBase(const _CSRIteratorBase<Other>& o) : cur(o.cur.m, o.cur.row, o.cur.index) {}
    _CSRIteratorBase(const _CSRIteratorBase&) = default;
    _CSRIteratorBase& operator=(const _CSRIteratorBase&) = default;
    bool operator==(const _CSRIteratorBase& o) const noexcept { return cur == o.cur; }
    bool operator!=(const _CSRIteratorBase& o) const noexcept { return !(cur == o.cur); }
    reference operator*() const { return const_cast<reference>(cur); }
    pointer operator->() const { return const_cast<pointer>(&cur); }
protected:
    int rowStart(int r) const { return cur.m->start[r]; }
    int rowCount() const { return cur.m->getDenseRowCount(); }
    int& curRow() { return cur.row; }
    int& curIndex() { return cur.index; }
    CSRElement cur;
};

// all stored elements, rows ascending, empty rows skipped
template <typename MatrixPtrT>
class CSRIterator : public _CSRIteratorBase<MatrixPtrT> {
    using Base = _CSRIteratorBase<MatrixPtrT>;
public:
    CSRIterator(MatrixPtrT m, int row, int index) noexcept : Base(m, row, index) {}
    template <typename Other, typename = std::enable_if_t<!std::is_same_v<Other, MatrixPtrT> && (!is_ptr_to_const_t<Other> || is_ptr_to_const_t<MatrixPtrT>)>>
    CSRIterator(const CSRIterator<Other>& o) : Base(o) {}
    CSRIterator& operator++() noexcept {
        const int next = ++this->curIndex();
        int r = this->curRow();
        while (r < this->rowCount() && next == this->rowStart(r + 1)) ++r;
        this->curRow() = r;
        return *this;
    }
    CSRIterator operator++(int) noexcept { CSRIterator t = *this; ++(*this); return t; }
};

// the elements of one row
template <typename MatrixPtrT>
class CSRRowIterator : public _CSRIteratorBase<MatrixPtrT> {
    using Base = _CSRIteratorBase<MatrixPtrT>;
public:
    CSRRowIterator(MatrixPtrT m, int row, int index) noexcept : Base(m, row, index) {}
    template <typename Other, typename = std::enable_if_t<!std::is_same_v<Other, MatrixPtrT> && (!is_ptr_to_const_t<Other> || is_ptr_to_const_t<MatrixPtrT>)>>
    CSRRowIterator(const CSRRowIterator<Other>& o) : Base(o) {}
    CSRRowIterator& operator++() noexcept {
        const int next = ++this->curIndex();
        if (next == this->rowStart(this->curRow() + 1)) ++this->curRow();
        return *this;
    }
    CSRRowIterator operator++(int) noexcept { CSRRowIterator t = *this; ++(*this); return t; }
};

enum class SolverPreconditioner {
    NONE,
    SYMMETRIC_GAUS_SEIDEL,
    ILU0,
    JACOBI,                                          // extension (not in the reference): diagonal preconditioner
    SYMMETRIC_GAUSS_SEIDEL = SYMMETRIC_GAUS_SEIDEL   // README spelling
};

// ---------------------------------------------------------------------------------------------------------------
// CSRMatrix<T> (ref H:1008-1651): host arrays are the source of truth for the element API; a device mirror is
// created on first use by a hot-path call and refreshed when the host side was mutated.
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
class CSRMatrix {
public:
    using Iterator = CSRIterator<CSRMatrix<T>*>;
    using ConstIterator = CSRIterator<const CSRMatrix<T>*>;
    using RowIterator = CSRRowIterator<CSRMatrix<T>*>;
    using ConstRowIterator = CSRRowIterator<const CSRMatrix<T>*>;
    using value_type = T;
    template <typename> friend class _CSRIteratorBase;

    CSRMatrix() noexcept = default;
    CSRMatrix(const TripletMatrix<T>& triplet) noexcept { init(triplet); }
    CSRMatrix(const CSRMatrix&) = delete;
    CSRMatrix& operator=(const CSRMatrix&) = delete;
    CSRMatrix(CSRMatrix&& o) noexcept { moveFrom(o); }
    CSRMatrix& operator=(CSRMatrix&& o) noexcept {
        if (this != &o) { releaseDevice(); moveFrom(o); }
        return *this;
    }
    ~CSRMatrix() { releaseDevice(); }

    // triplet -> CSR: count per row, prefix sum, scatter in the map's (row, col) order (ref H:1326-1349, 1606-1641)
    int init(const TripletMatrix<T>& triplet) noexcept {
        releaseDevice();
        denseRowCount = triplet.getDenseRowCount();
        denseColCount = triplet.getDenseColCount();
        const int nnz = triplet.getNonZeroCount();
        values.reset(new (std::nothrow) T[nnz > 0 ? nnz : 1]);
        positions.reset(new (std::nothrow) int[nnz > 0 ? nnz : 1]);
        start.reset(new (std::nothrow) int[denseRowCount + 1]);
        if (!values || !positions || !start) return 1;
        std::fill_n(start.get(), denseRowCount + 1, 0);
        for (const auto& el : triplet) start[el.getRow() + 1]++;
        firstActiveStart = denseRowCount;
        for (int r = 0; r < denseRowCount; ++r) {
            start[r + 1] += start[r];
            if (firstActiveStart == denseRowCount && start[r + 1] != 0) firstActiveStart = r;
        }
        int k = 0;
        for (const auto& el : triplet) {       // ordered map: rows ascending, columns ascending inside a row
            positions[k] = el.getCol();
            values[k] = el.getValue();
            ++k;
        }
        return 0;
    }
    // additive: direct CSR ingest (arrays are copied)
    int init(int rows, int cols, const int* startIn, const int* positionsIn, const T* valuesIn) noexcept {
        releaseDevice();
        denseRowCount = rows; denseColCount = cols;
        const int nnz = startIn[rows];
        values.reset(new (std::nothrow) T[nnz > 0 ? nnz : 1]);
        positions.reset(new (std::nothrow) int[nnz > 0 ? nnz : 1]);
        start.reset(new (std::nothrow) int[rows + 1]);
        if (!values || !positions || !start) return 1;
        std::copy_n(startIn, rows + 1, start.get());
        std::copy_n(positionsIn, nnz, positions.get());
        std::copy_n(valuesIn, nnz, values.get());
        firstActiveStart = rows;
        for (int r = 0; r < rows; ++r) if (start[r + 1] != 0) { firstActiveStart = r; break; }
        return 0;
    }

    int getNonZeroCount() const noexcept { return start ? start[denseRowCount] : 0; }
    int getDenseRowCount() const noexcept { return denseRowCount; }
    int getDenseColCount() const noexcept { return denseColCount; }
    bool hasSameNonZeroPattern(const CSRMatrix& o) {
        if (denseRowCount != o.denseRowCount || denseColCount != o.denseColCount) return false;
        const int nnz = getNonZeroCount();
        if (nnz != o.getNonZeroCount()) return false;
        return std::equal(start.get(), start.get() + denseRowCount, o.start.get()) &&
               std::equal(positions.get(), positions.get() + nnz, o.positions.get());
    }

    Iterator begin() noexcept { return Iterator(this, firstActiveStart, 0); }
    Iterator end() noexcept { return Iterator(this, denseRowCount, getNonZeroCount()); }
    ConstIterator begin() const noexcept { return cbegin(); }
    ConstIterator end() const noexcept { return cend(); }
    ConstIterator cbegin() const noexcept { return ConstIterator(this, firstActiveStart, 0); }
    ConstIterator cend() const noexcept { return ConstIterator(this, denseRowCount, getNonZeroCount()); }

    RowIterator rowBegin(const int i) noexcept { return RowIterator(this, i, start[i]); }
    RowIterator rowEnd(const int i) noexcept { return start[i] == start[i + 1] ? rowBegin(i) : RowIterator(this, i + 1, start[i + 1]); }
    ConstRowIterator rowBegin(const int i) const noexcept { return ConstRowIterator(this, i, start[i]); }
    ConstRowIterator rowEnd(const int i) const noexcept { return start[i] == start[i + 1] ? rowBegin(i) : ConstRowIterator(this, i + 1, start[i + 1]); }
    ConstRowIterator crowBegin(const int i) const noexcept { return rowBegin(i); }
    ConstRowIterator crowEnd(const int i) const noexcept { return rowEnd(i); }

    // ---- SpMV: the hot path (ref H:1458-1515) ----
    void rMult(const T* const mult, T* const out) const noexcept {
        b200::requireFloat<T>();
        assert(mult != out);
        spmv(SMM_OP_ASSIGN, nullptr, mult, out);
    }
    void rMultAdd(const T* const lhs, const T* const mult, T* const out) const noexcept {
        b200::requireFloat<T>();
        spmv(SMM_OP_ADD, lhs, mult, out);
    }
    void rMultSub(const T* const lhs, const T* const mult, T* const out) const noexcept {
        b200::requireFloat<T>();
        spmv(SMM_OP_SUB, lhs, mult, out);
    }

    // ---- host-side arithmetic and element access (ref H:1525-1604); each marks the device mirror stale ----
    void operator*=(const T scalar) { for (int i = 0, n = getNonZeroCount(); i < n; ++i) values[i] *= scalar; touchValues(); }
    void inplaceAdd(const CSRMatrix& o) {
        assert(hasSameNonZeroPattern(o));
        for (int i = 0, n = getNonZeroCount(); i < n; ++i) values[i] += o.values[i];
        touchValues();
    }
    void inplaceSubtract(const CSRMatrix& o) {
        assert(hasSameNonZeroPattern(o));
        for (int i = 0, n = getNonZeroCount(); i < n; ++i) values[i] -= o.values[i];
        touchValues();
    }
    bool updateEntry(const int row, const int col, const T newValue) {
        const int i = find(row, col);
        if (i < 0) return false;
        values[i] = newValue;
        touchValues();
        return true;
    }
    T getValue(const int row, const int col) const { const int i = find(row, col); return i < 0 ? T(0) : values[i]; }
    void zeroValues() { std::fill_n(values.get(), getNonZeroCount(), T(0)); touchValues(); }
    bool addEntry(const int row, const int col, const T value) {
        const int i = find(row, col);
        if (i < 0) return false;
        values[i] += value;
        touchValues();
        return true;
    }

    // ---- preconditioners (ref H:1165-1241) ----
    class IDPreconditioner {
    public:
        int apply(const T*, T*) const noexcept { return 0; }
    };

    class SGSPreconditioner {
    public:
        SGSPreconditioner(const CSRMatrix& matrix) noexcept : m(matrix) {}
        SGSPreconditioner(const SGSPreconditioner&) = delete;
        SGSPreconditioner& operator=(const SGSPreconditioner&) = delete;
        SGSPreconditioner(SGSPreconditioner&& o) noexcept : m(o.m), handle(o.handle), stamp(o.stamp) { o.handle = nullptr; }
        ~SGSPreconditioner() { if (handle) smm_precond_destroy(handle); }
        // (D+L) D^-1 (D+U) x = rhs by a forward and a backward sweep on the GPU; returns the reference's code
        int apply(const T* rhs, T* x) const noexcept {
            b200::requireFloat<T>();
            assert(rhs != x);
            int rc = 0;
            b200::check(smm_precond_apply(device(), rhs, x, &rc), "smm_precond_apply");
            return rc;
        }
        smm_precond_t* device() const {
            const smm_csr_t* a = m.device();
            if (!handle || stamp != m.structureStamp) {
                if (handle) smm_precond_destroy(handle);
                handle = nullptr;
                b200::check(smm_precond_sgs_create(a, &handle), "smm_precond_sgs_create");
                stamp = m.structureStamp;
            }
            return handle;
        }
        const CSRMatrix& matrix() const { return m; }
    private:
        const CSRMatrix& m;
        mutable smm_precond_t* handle = nullptr;
        mutable unsigned long long stamp = 0;
    };

    // EXTENSION (not in the reference): diagonal preconditioner, apply(rhs, x) is x_i = rhs_i / a_ii on the GPU, on the
    // matrix's current values.  getPreconditioner<SolverPreconditioner::JACOBI>() hands it out; BiCGStab accepts it.
    class JacobiPreconditioner {
    public:
        JacobiPreconditioner(const CSRMatrix& matrix) noexcept : m(matrix) {}
        JacobiPreconditioner(const JacobiPreconditioner&) = delete;
        JacobiPreconditioner& operator=(const JacobiPreconditioner&) = delete;
        JacobiPreconditioner(JacobiPreconditioner&& o) noexcept : m(o.m), handle(o.handle), stamp(o.stamp) { o.handle = nullptr; }
        ~JacobiPreconditioner() { if (handle) smm_precond_destroy(handle); }
        int apply(const T* rhs, T* x) const noexcept {
            b200::requireFloat<T>();
            assert(rhs != x);
            int rc = 0;
            b200::check(smm_precond_apply(device(), rhs, x, &rc), "smm_precond_apply");
            return rc;
        }
        smm_precond_t* device() const {
            const smm_csr_t* a = m.device();
            if (!handle || stamp != m.structureStamp) {
                if (handle) smm_precond_destroy(handle);
                handle = nullptr;
                b200::check(smm_precond_jacobi_create(a, &handle), "smm_precond_jacobi_create");
                stamp = m.structureStamp;
            }
            return handle;
        }
        const CSRMatrix& matrix() const { return m; }
    private:
        const CSRMatrix& m;
        mutable smm_precond_t* handle = nullptr;
        mutable unsigned long long stamp = 0;
    };

    // EXTENSION.  ILU(0) is dead code in the reference (factorize() can only fail, apply() is never defined, the
    // factory returns void for it; ref H:1188-1212, 1715-1790).  Here the type works: validate() performs the
    // zero-fill LU those lines describe on the host (0 ok, 1 unusable structure, 2 pivot not > 1e-6), apply() runs
    // L y = rhs, U x = y on the GPU, and BiCGStab accepts the object like any other preconditioner.
    class ILU0Preconditioner {
    public:
        ILU0Preconditioner(const CSRMatrix& matrix) noexcept : m(matrix) {}
        ILU0Preconditioner(const ILU0Preconditioner&) = delete;
        ILU0Preconditioner& operator=(const ILU0Preconditioner&) = delete;
        ILU0Preconditioner(ILU0Preconditioner&& o) noexcept : m(o.m), handle(o.handle), code(o.code) { o.handle = nullptr; }
        ~ILU0Preconditioner() { if (handle) smm_precond_destroy(handle); }
        int validate() noexcept {
            b200::requireFloat<T>();
            if (handle) { smm_precond_destroy(handle); handle = nullptr; }
            b200::check(smm_precond_ilu0_create(m.device(), &code, &handle), "smm_precond_ilu0_create");
            return code;
        }
        int apply(const T* rhs, T* x) const noexcept {
            b200::requireFloat<T>();
            if (!handle) return 1;
            int rc = 0;
            b200::check(smm_precond_apply(handle, rhs, x, &rc), "smm_precond_apply");
            return rc;
        }
        const smm_precond_t* device() const { return handle; }
        const CSRMatrix& matrix() const { return m; }
    private:
        const CSRMatrix& m;
        smm_precond_t* handle = nullptr;
        int code = 1;
    };

    // Zero-fill incomplete Cholesky (ref H:1214-1235): IC0Preconditioner M(m); M.init(); M.apply(rhs, x) and the
    // ConjugateGradient overload taking it.  init() factorises on the host (set-up), the solves run on the GPU.
    class IC0Preconditioner {
    public:
        IC0Preconditioner(const CSRMatrix& matrix) noexcept : m(matrix) {}
        IC0Preconditioner(const IC0Preconditioner&) = delete;
        IC0Preconditioner& operator=(const IC0Preconditioner&) = delete;
        IC0Preconditioner(IC0Preconditioner&& o) noexcept : m(o.m), handle(o.handle) { o.handle = nullptr; }
        ~IC0Preconditioner() { if (handle) smm_precond_destroy(handle); }
        int init() noexcept {
            b200::requireFloat<T>();
            if (handle) { smm_precond_destroy(handle); handle = nullptr; }
            int rc = 0;
            b200::check(smm_precond_ic0_create(m.device(), &rc, &handle), "smm_precond_ic0_create");
            return rc;
        }
        int apply(const T* rhs, T* x) const noexcept {
            b200::requireFloat<T>();
            int rc = 0;
            b200::check(smm_precond_apply(handle, rhs, x, &rc), "smm_precond_apply");
            return rc;
        }
        const smm_precond_t* device() const { return handle; }
        const CSRMatrix& matrix() const { return m; }
    private:
        const CSRMatrix& m;
        smm_precond_t* handle = nullptr;
    };

    template <SolverPreconditioner precond>
    decltype(auto) getPreconditioner() const noexcept {
        if constexpr (precond == SolverPreconditioner::NONE) return IDPreconditioner();
        else if constexpr (precond == SolverPreconditioner::SYMMETRIC_GAUS_SEIDEL) return SGSPreconditioner(*this);
        else if constexpr (precond == SolverPreconditioner::JACOBI) return JacobiPreconditioner(*this);
        else {                                          // extension: the reference's factory has no ILU0 branch (returns void)
            ILU0Preconditioner M(*this);
            M.validate();
            return M;
        }
    }

    // ---- device mirror ----
    smm_csr_t* device() const {
        b200::requireFloat<T>();
        if (!dev) {
            static const int zero = 0;
            b200::check(smm_csr_create(denseRowCount, denseColCount, start ? start.get() : &zero, positions.get(), values.get(), &dev), "smm_csr_create");
            valuesStale = false;
        } else if (valuesStale) {
            b200::check(smm_csr_update_values(dev, values.get()), "smm_csr_update_values");
            valuesStale = false;
        }
        return dev;
    }
    void touchValues() const noexcept { valuesStale = true; groupStale = true; }

    // ---- the same matrix row-partitioned over b200::devices() GPUs (created on first use) ----
    bool multiDevice() const noexcept { return b200::devices() > 1 && denseRowCount == denseColCount && denseRowCount > 0; }
    smm_group_t* group() const {
        b200::requireFloat<T>();
        const int want = b200::devices();
        // the reference-order reduction modes need row blocks that are nodes of the reference's reduction tree (H:308-320)
        const int part = b200::options().reduction_mode == SMM_REDUCE_FAST ? 0 : 1;
        if (grp && (grpDevices != want || grpPartition != part)) { smm_group_destroy(grp); grp = nullptr; }
        if (!grp) {
            b200::check(smm_group_create(denseRowCount, denseColCount, start.get(), positions.get(), values.get(), want, nullptr, part, &grp), "smm_group_create");
            grpDevices = want; grpPartition = part; groupStale = false;
        } else if (groupStale) {
            b200::check(smm_group_update_values(grp, values.get()), "smm_group_update_values");
            groupStale = false;
        }
        return grp;
    }

private:
    void spmv(int op, const T* lhs, const T* mult, T* out) const {
        if (multiDevice()) b200::check(smm_group_spmv(group(), op, lhs, mult, out), "smm_group_spmv");
        else b200::check(smm_spmv(device(), op, lhs, mult, out), "smm_spmv");
    }
    int find(const int row, const int col) const {     // binary search: columns ascend inside a row
        assert(row >= 0 && row < denseRowCount && col >= 0 && col < denseColCount);
        const int* b = positions.get() + start[row];
        const int* e = positions.get() + start[row + 1];
        const int* it = std::lower_bound(b, e, col);
        return (it != e && *it == col) ? static_cast<int>(it - positions.get()) : -1;
    }
    void releaseDevice() noexcept {
        if (dev) smm_csr_destroy(dev);
        dev = nullptr;
        if (grp) smm_group_destroy(grp);
        grp = nullptr;
        ++structureStamp;
    }
    void moveFrom(CSRMatrix& o) noexcept {
        values = std::move(o.values); positions = std::move(o.positions); start = std::move(o.start);
        denseRowCount = o.denseRowCount; denseColCount = o.denseColCount; firstActiveStart = o.firstActiveStart;
        dev = o.dev; valuesStale = o.valuesStale; structureStamp = o.structureStamp + 1;
        grp = o.grp; grpDevices = o.grpDevices; grpPartition = o.grpPartition; groupStale = o.groupStale;
        o.dev = nullptr; o.grp = nullptr; o.denseRowCount = o.denseColCount = 0; o.firstActiveStart = 0;
    }

    std::unique_ptr<T[]> values;        // [nnz]
    std::unique_ptr<int[]> positions;   // [nnz] column of each value, ascending inside a row
    std::unique_ptr<int[]> start;       // [rows + 1]
    int denseRowCount = 0;
    int denseColCount = 0;
    int firstActiveStart = 0;           // first non-empty row, or rows
    mutable smm_csr_t* dev = nullptr;
    mutable bool valuesStale = false;
    mutable smm_group_t* grp = nullptr;   // row blocks on b200::devices() GPUs
    mutable int grpDevices = 0, grpPartition = 0;
    mutable bool groupStale = false;
    mutable unsigned long long structureStamp = 1;
};

// ---------------------------------------------------------------------------------------------------------------
// dense text / dense array helpers (ref H:1930-2008); debugging formats, host only
// ---------------------------------------------------------------------------------------------------------------
template <typename T>
inline void saveDenseText(const char* filepath, const CSRMatrix<T>& m) {
    std::ofstream f(filepath);
    if (!f.is_open()) return;
    f << std::fixed << std::setprecision(6);
    const int rows = m.getDenseRowCount(), cols = m.getDenseColCount();
    f << rows << " " << cols << "\n{\n";
    std::vector<T> dense(static_cast<std::size_t>(cols));
    for (int r = 0; r < rows; ++r) {
        std::fill(dense.begin(), dense.end(), T(0));
        std::vector<char> stored(static_cast<std::size_t>(cols), 0);
        for (auto it = m.rowBegin(r); it != m.rowEnd(r); ++it) { dense[it->getCol()] = it->getValue(); stored[it->getCol()] = 1; }
        f << "{";
        for (int c = 0; c < cols; ++c) {
            if (stored[c]) f << dense[c]; else f << "0";
            if (c + 1 < cols) f << ",";
        }
        f << "}";
        if (r + 1 < rows) f << ",";
        f << "\n";
    }
    f << "}";
}

template <typename CompressedMatrixFormat>
inline void toLinearDenseRowMajor(const CompressedMatrixFormat& compressed, typename CompressedMatrixFormat::value_type* out) noexcept {
    const std::int64_t cols = compressed.getDenseColCount();
    for (const auto& el : compressed) out[el.getRow() * cols + el.getCol()] = el.getValue();
}

// ---------------------------------------------------------------------------------------------------------------
// Krylov solvers (ref H:2010-2398): thin forwards to the device-resident drivers
// ---------------------------------------------------------------------------------------------------------------
enum class SolverStatus { SUCCESS = 0, DIVERGED, MAX_ITERATIONS_REACHED };

template <typename T>
inline SolverStatus BiCGSymmetric(const CSRMatrix<T>& a, T* b, T* x, int maxIterations, T eps) {
    b200::requireFloat<T>();
    smm_solve_info info;
    if (a.multiDevice()) b200::check(smm_group_solve(a.group(), 1, b, nullptr, x, maxIterations, eps, &b200::options(), &info), "smm_group_solve");
    else b200::check(smm_solve_bicgsym(a.device(), b, x, maxIterations, eps, &b200::options(), &info), "smm_solve_bicgsym");
    b200::record(info);
    return static_cast<SolverStatus>(info.status);
}

template <typename T>
inline SolverStatus ConjugateGradientSquared(const CSRMatrix<T>& a, T* b, T* x, int maxIterations, T eps) {
    b200::requireFloat<T>();
    smm_solve_info info;
    if (a.multiDevice()) b200::check(smm_group_solve(a.group(), 2, b, nullptr, x, maxIterations, eps, &b200::options(), &info), "smm_group_solve");
    else b200::check(smm_solve_cgs(a.device(), b, x, maxIterations, eps, &b200::options(), &info), "smm_solve_cgs");
    b200::record(info);
    return static_cast<SolverStatus>(info.status);
}
template <typename T>
inline SolverStatus ConjugateGradientSqared(const CSRMatrix<T>& a, T* b, T* x, int maxIterations, T eps) {   // README spelling
    return ConjugateGradientSquared<T>(a, b, x, maxIterations, eps);
}

template <typename Preconditioner, typename T>
inline SolverStatus BiCGStab(const CSRMatrix<T>& a, T* b, T* x, int maxIterations, T eps, const Preconditioner& preconditioner) {
    b200::requireFloat<T>();
    using Id = typename CSRMatrix<T>::IDPreconditioner;
    using Sgs = typename CSRMatrix<T>::SGSPreconditioner;
    using Ic0 = typename CSRMatrix<T>::IC0Preconditioner;
    using Ilu0 = typename CSRMatrix<T>::ILU0Preconditioner;
    using Jac = typename CSRMatrix<T>::JacobiPreconditioner;
    // the reference's template takes any object with apply(rhs, x) (H:2191-2199); on this path the object must be one
    // whose apply lives on the GPU
    static_assert(std::is_same_v<Preconditioner, Id> || std::is_same_v<Preconditioner, Sgs> || std::is_same_v<Preconditioner, Ic0> ||
                      std::is_same_v<Preconditioner, Ilu0> || std::is_same_v<Preconditioner, Jac>,
                  "BiCGStab on the B200 path takes the preconditioner classes of CSRMatrix (ID, SGS, IC0, ILU0, Jacobi)");
    const smm_precond_t* p = nullptr;
    if constexpr (!std::is_same_v<Preconditioner, Id>) {
        assert(&preconditioner.matrix() == &a);
        p = preconditioner.device();
        if (!p) {
            std::fprintf(stderr, "sparse_matrix_math (B200): BiCGStab: the preconditioner has not been initialised (init() / validate())\n");
            std::abort();
        }
    }
    smm_solve_info info;
    // without a preconditioner the solve shards over b200::devices() GPUs; the triangular sweeps of a preconditioner do not
    if (p == nullptr && a.multiDevice()) b200::check(smm_group_solve(a.group(), 3, b, nullptr, x, maxIterations, eps, &b200::options(), &info), "smm_group_solve");
    else b200::check(smm_solve_bicgstab(a.device(), p, b, x, maxIterations, eps, &b200::options(), &info), "smm_solve_bicgstab");
    b200::record(info);
    return static_cast<SolverStatus>(info.status);
}

template <typename T>
inline SolverStatus BiCGStab(const CSRMatrix<T>& a, T* b, T* x, int maxIterations, T eps) {
    return BiCGStab(a, b, x, maxIterations, eps, a.template getPreconditioner<SolverPreconditioner::NONE>());
}

template <typename T>
inline SolverStatus ConjugateGradient(const CSRMatrix<T>& a, const T* const b, const T* const x0, T* const x, int maxIterations, T eps) {
    b200::requireFloat<T>();
    smm_solve_info info;
    if (a.multiDevice()) b200::check(smm_group_solve(a.group(), 0, b, x0, x, maxIterations, eps, &b200::options(), &info), "smm_group_solve");
    else b200::check(smm_solve_cg(a.device(), b, x0, x, maxIterations, eps, &b200::options(), &info), "smm_solve_cg");
    b200::record(info);
    return static_cast<SolverStatus>(info.status);
}

// preconditioned CG with the IC(0) object (ref H:2414-2505)
template <typename T>
inline SolverStatus ConjugateGradient(const CSRMatrix<T>& a, const T* const b, const T* const x0, T* const x, int maxIterations, T eps,
                                      const typename CSRMatrix<T>::IC0Preconditioner& M) {
    b200::requireFloat<T>();
    smm_solve_info info;
    b200::check(smm_solve_cg_ic0(a.device(), M.device(), b, x0, x, maxIterations, eps, &b200::options(), &info), "smm_solve_cg_ic0");
    b200::record(info);
    return static_cast<SolverStatus>(info.status);
}

// ---------------------------------------------------------------------------------------------------------------
// file loaders (ref H:2507-2669); setup code, host only
// ---------------------------------------------------------------------------------------------------------------
enum class MatrixLoadStatus {
    SUCCESS = 0,
    FAILED_TO_OPEN_FILE,
    FAILED_TO_OPEN_FILE_UNKNOWN_FORMAT,
    FAILED_TO_PARSE_FILE,
    PARSE_ERROR_MMX_FILE_MISSING_BANNER,
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_TYPE,
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_FORMAT,
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE,
    PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE
};

namespace b200 {
// One parser behind both loaders.  extended == false: exactly the reference's loadMatrixMarketMatrix (H:2531-2609).
// extended == true (not in the reference): also `general` and `skew-symmetric` structure and the
// `pattern` field (every stored entry is 1).
template <typename T>
inline MatrixLoadStatus loadMatrixMarketImpl(const char* filepath, TripletMatrix<T>& out, bool extended) {
    std::ifstream in(filepath);
    if (!in.is_open()) return MatrixLoadStatus::FAILED_TO_OPEN_FILE;
    auto lowered = [&in]() {
        std::string w;
        in >> w;
        std::transform(w.begin(), w.end(), w.begin(), [](unsigned char c) { return static_cast<char>(std::tolower(c)); });
        return w;
    };
    std::string banner;
    in >> banner;
    if (banner != "%%MatrixMarket") return MatrixLoadStatus::PARSE_ERROR_MMX_FILE_MISSING_BANNER;
    if (lowered() != "matrix") return MatrixLoadStatus::PARSE_ERROR_MMX_FILE_UNSUPPORTED_TYPE;
    if (lowered() != "coordinate") return MatrixLoadStatus::PARSE_ERROR_MMX_FILE_UNSUPPORTED_FORMAT;
    const std::string elType = lowered();
    const bool pattern = extended && elType == "pattern";
    if (elType != "real" && elType != "integer" && !pattern) return MatrixLoadStatus::PARSE_ERROR_MMX_FILE_UNSUPPORTED_EL_TYPE;
    const std::string structure = lowered();
    const bool general = extended && structure == "general";
    const bool skew = extended && structure == "skew-symmetric";
    if (structure != "symmetric" && !general && !skew) return MatrixLoadStatus::PARSE_ERROR_MMX_FILE_UNSUPPORTED_STRUCTURE;
    const auto skipLine = [&in]() { in.ignore(std::numeric_limits<std::streamsize>::max(), '\n'); };
    while (in.peek() == '%' || std::isspace(in.peek())) skipLine();
    int rows = 0, cols = 0, nnz = 0;
    in >> rows >> cols >> nnz;
    if (in.fail()) return MatrixLoadStatus::FAILED_TO_PARSE_FILE;
    out.init(rows, cols, nnz);
    if (extended && nnz == 0) return MatrixLoadStatus::SUCCESS;   // the reference's loop body runs at least once (H:2588)
    while (!in.eof()) {
        int r = 0, c = 0;
        T v = T(1);
        in >> r >> c;
        if (!pattern) in >> v;
        if (in.fail()) return MatrixLoadStatus::FAILED_TO_PARSE_FILE;
        out.addEntry(r - 1, c - 1, v);
        if (!general && r != c) out.addEntry(c - 1, r - 1, skew ? -v : v);
        while (std::isspace(in.peek())) skipLine();
    }
    return MatrixLoadStatus::SUCCESS;
}
}  // namespace b200

// coordinate x {real, integer} x symmetric; off-diagonals mirrored; explicit zeros kept
template <typename T>
inline MatrixLoadStatus loadMatrixMarketMatrix(const char* filepath, TripletMatrix<T>& out) {
    return b200::loadMatrixMarketImpl(filepath, out, false);
}

// "rows cols\n{{a,b,..},\n{..}}" as written by saveDenseText
template <typename T>
inline MatrixLoadStatus loadSMMDTMatrix(const char* filepath, TripletMatrix<T>& out) {
    std::ifstream in(filepath);
    if (!in.is_open()) return MatrixLoadStatus::FAILED_TO_OPEN_FILE;
    int rows = 0, cols = 0;
    in >> rows >> cols;
    if (in.fail()) return MatrixLoadStatus::FAILED_TO_PARSE_FILE;
    out.init(rows, cols, 0);
    const auto skipTo = [&in](char ch) { in.ignore(std::numeric_limits<std::streamsize>::max(), ch); };
    skipTo('{');
    for (int r = 0; r < rows; ++r) {
        skipTo('{');
        for (int c = 0; c < cols; ++c) {
            T v{};
            in >> v;
            if (in.fail()) return MatrixLoadStatus::FAILED_TO_PARSE_FILE;
            if (v != 0) out.addEntry(r, c, v);
            in.ignore(1, ',');
        }
        skipTo('\n');
    }
    // the closing line: a file that ends with its last row (no trailing newline, or truncated there) has already hit the
    // end of the stream and fails here, as in the reference (H:2640-2643)
    skipTo('\n');
    if (in.fail()) return MatrixLoadStatus::FAILED_TO_PARSE_FILE;
    return MatrixLoadStatus::SUCCESS;
}

template <typename T>
inline MatrixLoadStatus loadMatrix(const char* filepath, TripletMatrix<T>& out) {
    const char* dot = std::strrchr(filepath, '.');
    const std::string ext = dot ? dot + 1 : "";
    if (ext == "mtx") return loadMatrixMarketMatrix(filepath, out);
    if (ext == "smmdt") return loadSMMDTMatrix(filepath, out);
    return MatrixLoadStatus::FAILED_TO_OPEN_FILE_UNKNOWN_FORMAT;
}

template <typename T>
inline MatrixLoadStatus loadMatrix(const char* filepath, CSRMatrix<T>& out) {
    TripletMatrix<T> triplet;
    const MatrixLoadStatus status = loadMatrix(filepath, triplet);
    if (status != MatrixLoadStatus::SUCCESS) return status;
    out.init(triplet);
    return MatrixLoadStatus::SUCCESS;
}

// ---------------------------------------------------------------------------------------------------------------
// EXTENSION (not in the reference): Matrix Market files the reference rejects -- `general` and
// `skew-symmetric` structure, `pattern` field -- so that non-symmetric SuiteSparse matrices can be fed to
// CGS / BiCGStab.  The reference-named loaders above keep the reference's behaviour, including its rejections.
// ---------------------------------------------------------------------------------------------------------------
namespace ext {
template <typename T>
inline MatrixLoadStatus loadMatrixMarketMatrix(const char* filepath, TripletMatrix<T>& out) {
    return b200::loadMatrixMarketImpl(filepath, out, true);
}

template <typename T>
inline MatrixLoadStatus loadMatrix(const char* filepath, TripletMatrix<T>& out) {
    const char* dot = std::strrchr(filepath, '.');
    const std::string e = dot ? dot + 1 : "";
    if (e == "mtx") return ext::loadMatrixMarketMatrix(filepath, out);
    if (e == "smmdt") return loadSMMDTMatrix(filepath, out);
    return MatrixLoadStatus::FAILED_TO_OPEN_FILE_UNKNOWN_FORMAT;
}

template <typename T>
inline MatrixLoadStatus loadMatrix(const char* filepath, CSRMatrix<T>& out) {
    TripletMatrix<T> triplet;
    const MatrixLoadStatus status = ext::loadMatrix(filepath, triplet);
    if (status != MatrixLoadStatus::SUCCESS) return status;
    out.init(triplet);
    return MatrixLoadStatus::SUCCESS;
}
}  // namespace ext

}  // namespace SMM
