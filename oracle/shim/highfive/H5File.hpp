// Minimal no-op HighFive shim (test infrastructure, NOT product code).
// The reference's HDF5 I/O (src/hdf_interface.hpp:12-14) needs HighFive 2.9 + HDF5, which are
// not installed in this image.  Every HDF call site in the reference is templated, so declaring
// the handful of HighFive names it mentions is enough to compile the interpolation path; every
// method throws if it is ever executed.
#pragma once
#include <string>
#include <vector>
#include <stdexcept>
#include <initializer_list>
#include <utility>
namespace HighFive {
[[noreturn]] inline void nohdf(){ throw std::runtime_error("HDF5 support not built (HighFive shim)"); }
enum class ObjectType { File, Group, UserDataType, DataSpace, Dataset, Attribute, Other };
class DataType { public: DataType() = default; };
template<class T> class AtomicType : public DataType { public: AtomicType() = default; };
template<class T> class EnumType : public DataType {
public:
  struct member_def {
    std::string name; T value;
    member_def(const char* n, T v): name(n), value(v) {}
    member_def(std::string n, T v): name(std::move(n)), value(v) {}
  };
  EnumType(std::initializer_list<member_def>) {}
};
class CompoundType : public DataType {
public:
  struct member_def { std::string name; DataType type; member_def(const char* n, DataType t): name(n), type(t) {} };
  CompoundType(std::initializer_list<member_def>) {}
};
template<class T> DataType create_datatype();
class Object { public: virtual ~Object() = default; };
class Attribute : public Object {
public:
  template<class T> void read(T&) const { nohdf(); }
  template<class T> void write(const T&) { nohdf(); }
};
class DataSet : public Object {
public:
  template<class T> void read(T&) const { nohdf(); }
  template<class T> void read(T*) const { nohdf(); }
  template<class T> void write(const T&) { nohdf(); }
  std::vector<size_t> getDimensions() const { nohdf(); }
  size_t getElementCount() const { nohdf(); }
  template<class T> Attribute createAttribute(const std::string&, const T&) { nohdf(); }
  Attribute getAttribute(const std::string&) const { nohdf(); }
  bool hasAttribute(const std::string&) const { nohdf(); }
};
class Group;
template<class D> class NodeOps : public Object {
public:
  bool exist(const std::string&) const { nohdf(); }
  void unlink(const std::string&) { nohdf(); }
  template<class... A> Group createGroup(const std::string&, A...);
  Group getGroup(const std::string&) const;
  template<class... A> DataSet createDataSet(const std::string&, A...) { nohdf(); }
  DataSet getDataSet(const std::string&) const { nohdf(); }
  ObjectType getObjectType(const std::string&) const { nohdf(); }
  template<class T> Attribute createAttribute(const std::string&, const T&) { nohdf(); }
  Attribute getAttribute(const std::string&) const { nohdf(); }
  bool hasAttribute(const std::string&) const { nohdf(); }
  std::vector<std::string> listObjectNames() const { nohdf(); }
};
class Group : public NodeOps<Group> {};
template<class D> template<class... A> Group NodeOps<D>::createGroup(const std::string&, A...) { nohdf(); }
template<class D> Group NodeOps<D>::getGroup(const std::string&) const { nohdf(); }
class File : public NodeOps<File> {
public:
  enum : unsigned { ReadOnly=0x00u, ReadWrite=0x01u, Truncate=0x02u, Excl=0x04u, Debug=0x08u, Create=0x10u,
                    Overwrite=Truncate, OpenOrCreate=ReadWrite|Create };
  File(const std::string&, unsigned = ReadOnly) { nohdf(); }
};
}
