// TEST INFRASTRUCTURE.  Boost.Filesystem is not installed here; the reference's VO app and CCameraRecord.h use
// path, exists, create_directories, parent_path, operator/ and string() only -- all of which C++17's
// std::filesystem has under the same names.
#ifndef PHOVO_SHIM_BOOST_FILESYSTEM_PATH_HPP_
#define PHOVO_SHIM_BOOST_FILESYSTEM_PATH_HPP_
#include <filesystem>
namespace boost { namespace filesystem = std::filesystem; }
#endif
