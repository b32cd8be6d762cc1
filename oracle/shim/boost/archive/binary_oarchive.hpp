// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  boost::archive::binary_oarchive / binary_iarchive as far as util/FileIO.cpp names them: objects can be "streamed" into
// them and nothing happens (binary serialisation of frames / tracks is not on the compiled path; the pose TEXT files of FileIO.cpp use plain std streams).
#pragma once
#include <istream>
#include <ostream>
namespace boost { namespace archive {
struct binary_oarchive {
  explicit binary_oarchive(std::ostream&, unsigned = 0) {}
  template <class T> binary_oarchive& operator<<(const T&) { return *this; }
  template <class T> binary_oarchive& operator&(const T&) { return *this; }
};
struct binary_iarchive {
  explicit binary_iarchive(std::istream&, unsigned = 0) {}
  template <class T> binary_iarchive& operator>>(T&) { return *this; }
  template <class T> binary_iarchive& operator&(T&) { return *this; }
};
}  }
