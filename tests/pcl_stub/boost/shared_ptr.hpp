// boost::shared_ptr stand-in (tests/pcl_stub/README.md)
#pragma once
#include <memory>
namespace boost {
using std::shared_ptr;
using std::make_shared;
using std::static_pointer_cast;
using std::dynamic_pointer_cast;
}  // namespace boost
