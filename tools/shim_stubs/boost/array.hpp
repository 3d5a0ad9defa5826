#pragma once
#include <algorithm>
#include <cstddef>
#include <stdexcept>
namespace boost {
// a distinct type (openravejson.h overloads on std::array AND boost::array), aggregate like the real one
template <typename T, std::size_t N>
struct array {
    T elems[N ? N : 1];
    typedef T value_type;
    typedef T* iterator;
    typedef const T* const_iterator;
    typedef std::size_t size_type;
    iterator begin() { return elems; }
    const_iterator begin() const { return elems; }
    iterator end() { return elems + N; }
    const_iterator end() const { return elems + N; }
    T& operator[](size_type i) { return elems[i]; }
    const T& operator[](size_type i) const { return elems[i]; }
    T& at(size_type i) { if (i >= N) throw std::out_of_range("array"); return elems[i]; }
    const T& at(size_type i) const { if (i >= N) throw std::out_of_range("array"); return elems[i]; }
    T& front() { return elems[0]; }
    const T& front() const { return elems[0]; }
    T& back() { return elems[N - 1]; }
    const T& back() const { return elems[N - 1]; }
    static size_type size() { return N; }
    static bool empty() { return N == 0; }
    T* data() { return elems; }
    const T* data() const { return elems; }
    void fill(const T& v) { std::fill_n(elems, N, v); }
    void assign(const T& v) { fill(v); }
};
template <typename T, std::size_t N> bool operator==(const array<T, N>& a, const array<T, N>& b) { return std::equal(a.begin(), a.end(), b.begin()); }
template <typename T, std::size_t N> bool operator!=(const array<T, N>& a, const array<T, N>& b) { return !(a == b); }
}
