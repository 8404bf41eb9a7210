// TEST INFRASTRUCTURE: stand-in for the boost::serialization names the reference's DBoW2 headers mention inside serialize() templates that are
// never instantiated here.
#pragma once
namespace boost { namespace serialization {
class access;
template <class Base, class Derived> Base& base_object(Derived& d) { return static_cast<Base&>(d); }
} }
