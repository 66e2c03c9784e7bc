"""SVDD-PM decoding CLI -- drop-in for the reference's ``decode_tweedie.py`` (a 10-line
variant of decode.py: required string flag ``--tweedie`` compared to "True" at
diffusion_gosai.py:1414, output suffix ``_tw`` at decode_tweedie.py:118)."""
import decode

if __name__ == '__main__':
  decode.run(decode.build_parser(tweedie=True).parse_args(), tweedie=True)
