"""Dev-time helper: read the reference's FORWARD VLC tables (ISO/IEC 13818-2 Annex B data, as held in
/root/reference/src/core/mp2v_luts.hpp) into plain Python dicts of bit-strings, so that
tests/test_vlc_tables.py can check this repo's own table file against them when the reference is
mounted.  Nothing here is used at run time and nothing is copied into the product."""
import re, os

REF = os.environ.get("MP2V_REFERENCE", "/root/reference")

def _body(src, name):
    m = re.search(r'\b%s\s*\[\d+\]\s*=\s*\{(.*?)\n\};' % re.escape(name), src, re.S)
    return m.group(1)

def _strip_comments(s):
    return re.sub(r'//[^\n]*', '', s)

def load():
    src = open(os.path.join(REF, "src/core/mp2v_luts.hpp")).read()
    out = {}
    def vlc_list(name):
        body = _strip_comments(_body(src, name))
        items = re.findall(r'\{\s*0b([01]+)\s*,\s*(\d+)\s*\}', body)
        return [format(int(b, 2), '0%db' % int(n)) if int(n) else '' for b, n in items]
    out['mba'] = vlc_list('macroblock_address_increment_to_vlc')      # index = increment (0 unused)
    out['cbp'] = vlc_list('coded_block_pattern_to_vlc')               # index = cbp value
    out['motion'] = vlc_list('motion_code_to_vlc')                    # index = code + 16
    out['dc_luma'] = vlc_list('dct_size_luminance_to_vlc')
    out['dc_chroma'] = vlc_list('dct_size_chrominance_to_vlc')
    def coef_list(name):
        body = _strip_comments(_body(src, name))
        items = re.findall(r'\{\s*\{\s*0b([01]+)\s*,\s*(\d+)\s*\}\s*,\s*\{\s*(\d+)\s*,\s*(\d+)\s*\}\s*\}', body)
        return {(int(r), int(l)): format(int(b, 2), '0%db' % int(n)) for b, n, r, l in items}
    out['b14'] = coef_list('coeff_zero_vlc')
    out['b15'] = coef_list('coeff_one_vlc')
    def mbtype(name):
        body = _strip_comments(_body(src, name))
        items = re.findall(r'\{\s*\{\s*0b([01]+)\s*,\s*(\d+)\s*\}\s*,\s*0b([01]+)\s*\}', body)
        return {int(v, 2): format(int(b, 2), '0%db' % int(n)) for b, n, v in items}
    out['mbtype_i'] = mbtype('i_macroblock_type')
    out['mbtype_p'] = mbtype('p_macroblock_type')
    out['mbtype_b'] = mbtype('b_macroblock_type')
    return out

if __name__ == "__main__":
    t = load()
    for k, v in t.items():
        print(k, len(v))
    b14 = t['b14']
    # dump run/level tables sorted by (len, code) for eyeballing
    for name in ('b14', 'b15'):
        print(name)
        for (r, l), bits in sorted(t[name].items(), key=lambda kv: (len(kv[1]), kv[1])):
            print("   ", bits, r, l)
