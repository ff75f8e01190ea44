// Compile-only stand-in for yaml-cpp (absent from this image).  TEST INFRASTRUCTURE (oracle/ref_host): lets the
// reference's sources that PARSE configuration be compiled where they lie; the code paths the harness drives never
// touch a YAML::Node -- any use at run time aborts loudly.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <utility>
#include <vector>
#include <ostream>
namespace YAML {
[[noreturn]] inline void no_yaml(const char* what) { std::fprintf(stderr, "ref_host stub: YAML::%s needs yaml-cpp\n", what); std::abort(); }
namespace NodeType { enum value { Undefined, Null, Scalar, Sequence, Map }; }
class Node;
struct NodePair;
class NodeIter {
 public:
  NodeIter() {}
  NodePair* operator->() const;
  NodePair& operator*() const;
  NodeIter& operator++() { return *this; }
  NodeIter operator++(int) { return *this; }
  bool operator==(const NodeIter&) const { return true; }
  bool operator!=(const NodeIter&) const { return false; }
};
class Node {
 public:
  typedef NodeIter iterator;
  typedef NodeIter const_iterator;
  Node() {}
  template <class T> explicit Node(const T&) {}
  template <class K> Node operator[](const K&) const { no_yaml("operator[]"); }
  template <class T> T as() const { no_yaml("as"); }
  template <class T, class S> T as(const S&) const { no_yaml("as"); }
  template <class T> Node& operator=(const T&) { return *this; }
  explicit operator bool() const { return false; }
  bool operator!() const { return true; }
  bool IsDefined() const { return false; }
  bool IsNull() const { return true; }
  bool IsScalar() const { return false; }
  bool IsSequence() const { return false; }
  bool IsMap() const { return false; }
  NodeType::value Type() const { return NodeType::Undefined; }
  std::size_t size() const { return 0; }
  NodeIter begin() const { return NodeIter(); }
  NodeIter end() const { return NodeIter(); }
  std::string Scalar() const { return std::string(); }
  std::string Tag() const { return std::string(); }
  template <class T> void push_back(const T&) {}
  template <class K> bool remove(const K&) { return false; }
  void reset(const Node& = Node()) {}
  bool is(const Node&) const { return false; }
};
struct NodePair : Node { Node first, second; };
inline NodePair* NodeIter::operator->() const { no_yaml("iterator"); }
inline NodePair& NodeIter::operator*() const { no_yaml("iterator"); }
inline Node LoadFile(const std::string&) { no_yaml("LoadFile"); }
inline Node Load(const std::string&) { no_yaml("Load"); }
inline Node Load(const char*) { no_yaml("Load"); }
inline Node Clone(const Node&) { return Node(); }
inline std::string Dump(const Node&) { return std::string(); }
inline std::ostream& operator<<(std::ostream& o, const Node&) { return o; }
class Emitter {
 public:
  template <class T> Emitter& operator<<(const T&) { return *this; }
  const char* c_str() const { return ""; }
};
class Exception : public std::exception {};
class BadConversion : public Exception {};
class ParserException : public Exception {};
class InvalidNode : public Exception {};
class BadFile : public Exception {};
class RepresentationException : public Exception {};
template <class T> class TypedBadConversion : public BadConversion {};
template <class T> struct convert;
}  // namespace YAML
