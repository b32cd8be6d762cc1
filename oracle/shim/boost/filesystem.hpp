// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  boost::filesystem as far as the compiled reference files name it; the stand-in NEVER touches the file system:
// nothing exists, nothing is created or removed (the reference only uses it around debug / result files).
#pragma once
#include <string>
namespace boost { namespace filesystem {
struct path { std::string s; path(const std::string& p = "") : s(p) {} path(const char* p) : s(p) {} const std::string& string() const { return s; } };
inline bool exists(const path&) { return false; }
inline bool is_directory(const path&) { return false; }
inline bool create_directories(const path&) { return false; }
inline bool create_directory(const path&) { return false; }
inline unsigned long remove_all(const path&) { return 0; }
inline bool remove(const path&) { return false; }
}  }
namespace boost { namespace filesystem {
// an always-empty directory listing
struct file_status {};
struct directory_entry { filesystem::path p; const filesystem::path& path() const { return p; } file_status status() const { return file_status(); } };
struct directory_iterator {
  directory_iterator() {} explicit directory_iterator(const filesystem::path&) {}
  bool operator!=(const directory_iterator&) const { return false; } bool operator==(const directory_iterator&) const { return true; }
  directory_iterator& operator++() { return *this; } directory_entry operator*() const { return directory_entry(); } const directory_entry* operator->() const { static directory_entry e; return &e; }
};
inline bool is_regular_file(const file_status&) { return false; }
inline bool is_regular_file(const path&) { return false; }
}  }
