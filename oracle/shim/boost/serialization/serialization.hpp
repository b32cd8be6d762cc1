// ORACLE/shim: the names of boost::serialization that the reference's headers mention (friend declarations, free serialize() templates that are never
// instantiated here); nothing is serialised on the compiled path.
#pragma once
#include <cstddef>
namespace boost { namespace serialization {
class access {};
template <class T> struct array_wrapper { T* p; std::size_t n; };
template <class T> inline array_wrapper<T> make_array(T* p, std::size_t n) { return array_wrapper<T>{p, n}; }
template <class Archive, class T> inline void split_free(Archive&, T&, const unsigned int) {}
}  }
#define BOOST_SERIALIZATION_SPLIT_FREE(T)
#define BOOST_SERIALIZATION_SPLIT_MEMBER()
