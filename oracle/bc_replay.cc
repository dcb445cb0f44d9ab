// oracle/bc_replay.cc — TEST INFRASTRUCTURE ONLY.
// Oracle for the ORDER of the `-b` barcode file (print_barcodes, /root/reference/src/junctions/junctions_extractor.h:99-111).
// The reference keeps a std::unordered_map<string,int> per junction and copy-assigns it once per supporting read
// (junctions_extractor.cc:203-215); a copy keeps bucket count and node order, so the printed order equals inserting the
// junction's distinct barcodes, in first-seen order, into ONE map and iterating it.  This helper does exactly that with the
// toolchain's own libstdc++ (the order is a property of that library, not of regtools).
//   stdin : one line per junction: "barcode count barcode count ..." in first-seen order
//   stdout: "<n>\t<bc>:<count>,<bc>:<count>...\n" as print_barcodes writes it
#include <iostream>
#include <sstream>
#include <string>
#include <unordered_map>
int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::unordered_map<std::string, int> m;
        std::string bc; int c;
        while (is >> bc >> c) m.insert(std::pair<std::string, int>(bc, c));
        std::cout << m.size() << "\t";
        for (std::unordered_map<std::string, int>::const_iterator it = m.begin(); it != m.end(); ++it) {
            if (it != m.begin()) std::cout << ",";
            std::cout << it->first << ":" << it->second;
        }
        std::cout << std::endl;
    }
    return 0;
}
