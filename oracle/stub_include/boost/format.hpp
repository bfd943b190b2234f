// Stand-in (TEST INFRASTRUCTURE): boost::format is only used by the reference's LM debug print
// (lsq_registration_impl.hpp:203-212, off by default); arguments are swallowed, the format string is printed.
#ifndef DDLO_ORACLE_BOOST_FORMAT_STUB
#define DDLO_ORACLE_BOOST_FORMAT_STUB
#include <ostream>
#include <string>
namespace boost {
class format {
  std::string s_;

 public:
  explicit format(const char* s) : s_(s) {}
  template <class T>
  format& operator%(const T&) {
    return *this;
  }
  friend std::ostream& operator<<(std::ostream& os, const format& f) { return os << f.s_; }
};
}  // namespace boost
#endif
