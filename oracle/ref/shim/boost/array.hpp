/* Shim of boost::array (Boost 1.53 is not vendored with the reference): a fixed-size aggregate. */
#ifndef REF_SHIM_BOOST_ARRAY_HPP
#define REF_SHIM_BOOST_ARRAY_HPP
#include <cstddef>
namespace boost {
template <class T, std::size_t N>
struct array {
    T elems[N];
    T& operator[](std::size_t i) { return elems[i]; }
    const T& operator[](std::size_t i) const { return elems[i]; }
    static std::size_t size() { return N; }
};
}
#endif
