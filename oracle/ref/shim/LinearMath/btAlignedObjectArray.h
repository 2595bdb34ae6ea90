/* Shim (oracle/ref): btAlignedObjectArray as the reference uses it -- size / clear / push_back / operator[]. */
#ifndef REF_SHIM_BT_ALIGNED_OBJECT_ARRAY_H
#define REF_SHIM_BT_ALIGNED_OBJECT_ARRAY_H
#include <vector>
template <class T>
class btAlignedObjectArray {
public:
    int size() const { return (int)m_v.size(); }
    void clear() { m_v.clear(); }
    void push_back(const T& t) { m_v.push_back(t); }
    T& operator[](int i) { return m_v[i]; }
    const T& operator[](int i) const { return m_v[i]; }
private:
    std::vector<T> m_v;
};
#endif
