#pragma once
#include <array>
#include <complex>   // the real header pulls it in; mathextra.h relies on that
#include <cstddef>
#include <vector>
namespace boost {
template <std::size_t N> struct extent_gen_stub { std::array<std::size_t, N> e; extent_gen_stub<N + 1> operator[](std::size_t n) const { extent_gen_stub<N + 1> r; for (std::size_t i = 0; i < N; ++i) r.e[i] = e[i]; r.e[N] = n; return r; } };
template <> struct extent_gen_stub<0> { extent_gen_stub<1> operator[](std::size_t n) const { extent_gen_stub<1> r; r.e[0] = n; return r; } };
static const extent_gen_stub<0> extents = extent_gen_stub<0>();
template <typename T, std::size_t N>
class multi_array {
public:
    typedef T element;
    typedef std::size_t size_type;
    typedef std::size_t index;
    multi_array() { _shape.fill(0); }
    explicit multi_array(const extent_gen_stub<N>& e) { resize(e); }
    void resize(const extent_gen_stub<N>& e) { _shape = e.e; std::size_t n = 1; for (std::size_t s : _shape) n *= s; _data.assign(n, T()); }
    const size_type* shape() const { return _shape.data(); }
    size_type num_elements() const { return _data.size(); }
    size_type size() const { return _shape[0]; }
    T* data() { return _data.data(); }
    const T* data() const { return _data.data(); }
    T* origin() { return _data.data(); }
    const T* origin() const { return _data.data(); }
    struct row { T* p; T& operator[](std::size_t j) const { return p[j]; } };
    row operator[](std::size_t i) { std::size_t stride = 1; for (std::size_t k = 1; k < N; ++k) stride *= _shape[k]; return row{_data.data() + i * stride}; }
    const row operator[](std::size_t i) const { return const_cast<multi_array*>(this)->operator[](i); }
private:
    std::array<std::size_t, N> _shape;
    std::vector<T> _data;
};
}
