// Stand-in (TEST INFRASTRUCTURE): boost::shared_ptr as the reference's nanoflann wrapper uses it.
#ifndef DDLO_ORACLE_BOOST_SHARED_PTR_STUB
#define DDLO_ORACLE_BOOST_SHARED_PTR_STUB
#include <memory>
namespace boost {
template <class T>
using shared_ptr = std::shared_ptr<T>;
using std::make_shared;
}  // namespace boost
#endif
