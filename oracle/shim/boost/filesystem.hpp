// Test-infrastructure shim (oracle/Makefile, oracle/_ref only): lets UNMODIFIED reference sources compile without Boost.Filesystem (scan.h / basicScan.h name path, exists, last_write_time).
#pragma once
#include <ctime>
#include <filesystem>
#include <functional>
#include <chrono>
namespace boost { namespace filesystem {
using std::filesystem::path;
using std::filesystem::exists;
using std::filesystem::is_directory;
using std::filesystem::create_directory;
using std::filesystem::create_directories;
using std::filesystem::directory_iterator;
inline std::time_t last_write_time(const path& p) {
  auto t = std::filesystem::last_write_time(p);
  auto s = std::chrono::time_point_cast<std::chrono::system_clock::duration>(
      t - std::filesystem::file_time_type::clock::now() + std::chrono::system_clock::now());
  return std::chrono::system_clock::to_time_t(s);
}
} }
