// Stub of colmap/util/endian.h: little-endian host assumed (x86-64).
#pragma once
#include <cstring>
#include <istream>
#include <ostream>
#include <vector>
namespace colmap {
template <typename T> void ReadBinaryLittleEndian(std::istream* s, std::vector<T>* data) { s->read(reinterpret_cast<char*>(data->data()), data->size() * sizeof(T)); }
template <typename T> void WriteBinaryLittleEndian(std::ostream* s, const std::vector<T>& data) { s->write(reinterpret_cast<const char*>(data.data()), data.size() * sizeof(T)); }
}  // namespace colmap
