// Force-included (g++ -include) when compiling the UNMODIFIED reference with a
// modern g++: two-phase lookup cannot see apex_exp_template::operators from
// ContainerExp::operator+= (apex-tensor/apex_exp_template.h:328-379) unless the
// namespace is inline.  No reference source is edited.
namespace apex_exp_template { inline namespace operators {} }
