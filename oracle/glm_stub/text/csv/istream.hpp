/*
 * text-csv stand-in (TEST INFRASTRUCTURE ONLY) so the reference's unmodified
 * fileloader.cpp compiles into oracle/_ref.  The real package
 * (text-csv/latest@signal9/stable, conanfile.txt:4) is an un-vendored Conan
 * dependency; only the three operations load_lights uses are restated
 * (fileloader.cpp:123-133): construct from an istream, test for more input,
 * extract one comma/newline separated float.
 */
#ifndef TEXT_CSV_ISTREAM_STANDIN_HPP
#define TEXT_CSV_ISTREAM_STANDIN_HPP
#include <cstdlib>
#include <istream>
#include <string>

namespace text { namespace csv {
class csv_istream {
    std::istream &in_;
public:
    explicit csv_istream(std::istream &in) : in_(in) {}
    /* true while another field can be read (blank tail / EOF => false) */
    explicit operator bool() {
        while (in_.good()) {
            int c = in_.peek();
            if (c == '\n' || c == '\r' || c == ' ') { in_.get(); continue; }
            break;
        }
        return in_.good() && in_.peek() != std::char_traits<char>::eof();
    }
    csv_istream &operator>>(float &value) {
        std::string field;
        int c;
        while ((c = in_.get()) != std::char_traits<char>::eof() && c != ',' && c != '\n') field.push_back((char)c);
        value = std::strtof(field.c_str(), 0);
        return *this;
    }
};
}} // namespace text::csv
#endif
