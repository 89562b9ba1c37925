"""CPU oracle for the HoneyBADGER HMM segmentation path — TEST INFRASTRUCTURE ONLY.

Nothing under honeybadger_b200/ may import this package.  See hb_oracle.c for the parity status
("parity unpinned": the reference's arithmetic lives in HiddenMarkov / R nmath, absent here).
"""
