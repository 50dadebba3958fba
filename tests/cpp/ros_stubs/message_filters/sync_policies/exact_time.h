#pragma once
namespace message_filters { namespace sync_policies {
struct NullType {};
template <class M0, class M1, class M2 = NullType, class M3 = NullType> struct ExactTime { explicit ExactTime(unsigned queue_size); };
} }
