// Stand-in for <boost/program_options.hpp>: CommandLine.h only needs the types to exist (CommandLine.cpp is not
// part of the oracle build).  TEST INFRASTRUCTURE.
#ifndef PHASE_ORACLE_PROGRAM_OPTIONS_STUB
#define PHASE_ORACLE_PROGRAM_OPTIONS_STUB
#include <map>
#include <stdexcept>
#include <string>
namespace boost { namespace program_options {
template <class T> struct typed_value { typed_value *required() { return this; } };
template <class T> typed_value<T> *value() { static typed_value<T> v; return &v; }
struct options_description_easy_init {
  template <class V> options_description_easy_init &operator()(const char *, V *, const char *) { return *this; }
  options_description_easy_init &operator()(const char *, const char *) { return *this; }
};
struct options_description { options_description_easy_init add_options() { return options_description_easy_init(); } };
struct variable_value {
  template <class T> const T &as() const { throw std::logic_error("no command line in the oracle build"); }
};
struct variables_map {
  const variable_value &operator[](const std::string &) const { static variable_value v; return v; }
  std::size_t count(const std::string &) const { return 0; }
};
}}
#endif
