// ORACLE/shim — TEST INFRASTRUCTURE ONLY.  Stand-in for the part of boost::program_options that the reference's base/Config.h uses: an options_description
// filled through add_options()(name, value<T>(&target), help), parse_config_file() for "key = value" lines ('#' comments, blank lines), store() / notify().
// Values are converted with operator>> (bool: true / false / 1 / 0 / on / off / yes / no, like program_options); an unknown key throws, as the original does.
#pragma once
#include <algorithm>
#include <cctype>
#include <istream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace boost { namespace program_options {
struct value_semantic { virtual ~value_semantic() {} virtual void set(const std::string& s) const = 0; };
template <typename T> struct typed_value : value_semantic {
  T* target;
  explicit typed_value(T* t) : target(t) {}
  void set(const std::string& s) const override { std::istringstream is(s); T v; if (!(is >> v)) throw std::runtime_error("invalid option value '" + s + "'"); *target = v; }
};
template <> inline void typed_value<std::string>::set(const std::string& s) const { *target = s; }
template <> inline void typed_value<bool>::set(const std::string& s) const {
  std::string v = s; std::transform(v.begin(), v.end(), v.begin(), [](unsigned char c) { return (char)std::tolower(c); });
  if (v == "true" || v == "1" || v == "on" || v == "yes") *target = true;
  else if (v == "false" || v == "0" || v == "off" || v == "no") *target = false;
  else throw std::runtime_error("invalid bool value '" + s + "'");
}
template <typename T> inline typed_value<T>* value(T* t) { return new typed_value<T>(t); }
template <typename T> inline typed_value<T>* value() { static T sink; return new typed_value<T>(&sink); }

class options_description {
 public:
  std::map<std::string, std::shared_ptr<const value_semantic>> opts;
  explicit options_description(const std::string& = "") {}
  struct easy_init {
    options_description* d;
    easy_init& operator()(const char* name, const value_semantic* v, const char* = "") { d->opts[name] = std::shared_ptr<const value_semantic>(v); return *this; }
    easy_init& operator()(const char* name, const char* = "") { d->opts[name] = nullptr; return *this; }
  };
  easy_init add_options() { return easy_init{this}; }
};
struct parsed_options { std::vector<std::pair<std::string, std::string>> kv; const options_description* desc; };
class variables_map : public std::map<std::string, std::string> {
 public:
  const options_description* desc = nullptr;
  size_t count(const std::string& k) const { return std::map<std::string, std::string>::count(k); }
};
inline std::string trim_(const std::string& s) { size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n"); return a == std::string::npos ? "" : s.substr(a, b - a + 1); }
inline parsed_options parse_config_file(std::istream& in, const options_description& d, bool allow_unregistered = false) {
  parsed_options p; p.desc = &d;
  std::string line;
  while (std::getline(in, line)) {
    const size_t h = line.find('#'); if (h != std::string::npos) line = line.substr(0, h);
    line = trim_(line); if (line.empty() || line[0] == '[') continue;
    const size_t e = line.find('='); if (e == std::string::npos) throw std::runtime_error("invalid config line '" + line + "'");
    const std::string k = trim_(line.substr(0, e)), v = trim_(line.substr(e + 1));
    if (!d.opts.count(k)) { if (allow_unregistered) continue; throw std::runtime_error("unrecognised option '" + k + "'"); }
    p.kv.push_back(std::make_pair(k, v));
  }
  return p;
}
inline void store(const parsed_options& p, variables_map& vm) {
  vm.desc = p.desc;
  for (const auto& kv : p.kv) if (!vm.count(kv.first)) vm[kv.first] = kv.second;      // first occurrence wins
}
inline void notify(variables_map& vm) {
  if (!vm.desc) return;
  for (const auto& kv : vm) { auto it = vm.desc->opts.find(kv.first); if (it != vm.desc->opts.end() && it->second) it->second->set(kv.second); }
}
}  }
